B2D_ROI_IL=0 timeout 120 python scripts/bench_kernels.py il 2>&1 | tail -1
B2D_ROI_IL=1 timeout 120 python scripts/bench_kernels.py il 2>&1 | tail -1
