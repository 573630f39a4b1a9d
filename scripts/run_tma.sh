set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused or sampler or targets" 2>&1 | tail -15
timeout 120 python scripts/bench_kernels.py stages 2>&1 | tail -6
