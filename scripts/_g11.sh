mkdir -p gpurun_out/r11
python scripts/heis_probe.py 27 0 1 500 1000 2000 3000 5000 > gpurun_out/r11/probe.txt 2>&1
cat gpurun_out/r11/probe.txt
