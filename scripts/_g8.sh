mkdir -p gpurun_out/r8
CMPT_B200_TRACE_SLAB=1 python scripts/virtual_checks_debug.py 4 heisenberg_mf > gpurun_out/r8/v4.txt 2>&1
CMPT_B200_NO_FUSED_SLAB=1 python scripts/virtual_checks_debug.py 4 heisenberg_mf > gpurun_out/r8/v4_nofused.txt 2>&1
python scripts/heis_probe.py 27 0 -1 16 12 8 4 > gpurun_out/r8/probe.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:heis_apply -s 2 -c 2 -o gpurun_out/r8/heis_cluster python scripts/heis_probe.py 27 -1 > gpurun_out/r8/ncu.log 2>&1
tail -n 3 gpurun_out/r8/v4.txt gpurun_out/r8/v4_nofused.txt | cut -c1-300; cat gpurun_out/r8/probe.txt
