timeout 120 python scripts/bench_kernels.py stages 2>&1 | grep roi_align
timeout 300 python -m pytest tests -m gpu -q -k roi 2>&1 | tail -2
