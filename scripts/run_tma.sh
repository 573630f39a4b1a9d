set -x
export B2D_ROI_TMA_DEV=-7
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_roi_align_tma -s 4 -c 1 -o gpurun_out/prof_tma3 python scripts/bench_kernels.py tma > gpurun_out/ncu_tma.log 2>&1
echo rc=$?
