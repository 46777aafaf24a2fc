mkdir -p gpurun_out/r9
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "heisenberg" -x 2>&1 | tail -n 15 > gpurun_out/r9/heis_tests.txt
python scripts/heis_probe.py 27 0 1 > gpurun_out/r9/probe.txt 2>&1
python scripts/heis_probe.py 24 0 1 >> gpurun_out/r9/probe.txt 2>&1
timeout 120 python scripts/virtual_checks_debug.py 4 heisenberg_mf > gpurun_out/r9/v4.txt 2>&1
timeout 120 python scripts/virtual_checks_debug.py 8 heisenberg_mf > gpurun_out/r9/v8.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:heis_apply -s 2 -c 2 -o gpurun_out/r9/heis_siblings python scripts/heis_probe.py 27 1 > gpurun_out/r9/ncu.log 2>&1
tail -n 3 gpurun_out/r9/heis_tests.txt gpurun_out/r9/v4.txt gpurun_out/r9/v8.txt | cut -c1-300; cat gpurun_out/r9/probe.txt
