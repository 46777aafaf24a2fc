mkdir -p gpurun_out/r10
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "heisenberg" -x 2>&1 | tail -n 5 > gpurun_out/r10/heis_tests.txt
python scripts/heis_probe.py 27 0 1 > gpurun_out/r10/probe.txt 2>&1
python scripts/heis_probe.py 25 0 1 >> gpurun_out/r10/probe.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:heis_apply -s 2 -c 2 -o gpurun_out/r10/heis_siblings python scripts/heis_probe.py 27 1 > gpurun_out/r10/ncu.log 2>&1
tail -n 3 gpurun_out/r10/heis_tests.txt; cat gpurun_out/r10/probe.txt
