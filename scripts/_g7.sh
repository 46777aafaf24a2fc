mkdir -p gpurun_out/r7
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "heisenberg" -x 2>&1 | tail -15 > gpurun_out/r7/heis_tests.txt
python scripts/virtual_checks_debug.py 2 heisenberg_mf > gpurun_out/r7/v2.txt 2>&1
python scripts/virtual_checks_debug.py 8 heisenberg_mf > gpurun_out/r7/v8.txt 2>&1
python bench.py --config 5 --steps 3 --warmup 3 > gpurun_out/r7/bench5_cluster.json 2> gpurun_out/r7/bench5_cluster.err
CMPT_B200_HEIS_NO_CLUSTER=1 python bench.py --config 5 --steps 3 --warmup 3 > gpurun_out/r7/bench5_nocluster.json 2> gpurun_out/r7/bench5_nocluster.err
CMPT_B200_SPIN_TIMEOUT_S=8 python -m pytest tests/test_virtual_ranks.py tests/test_cpp_samples.py tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -12 > gpurun_out/r7/tests_rest.txt
tail -3 gpurun_out/r7/*.txt; cut -c1-600 gpurun_out/r7/*.json
