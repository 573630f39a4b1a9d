set -x
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_roi_align_tma -s 4 -c 1 -o gpurun_out/prof_tma python scripts/bench_kernels.py tma > gpurun_out/ncu_tma.log 2>&1
echo rc=$?
