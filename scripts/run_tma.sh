set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 120 python scripts/bench_kernels.py stages 2>&1 | tail -6
B2D_RPN_CHAINS=0 timeout 120 python scripts/bench_kernels.py stages 2>&1 | tail -6
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_chain.json 2> gpurun_out/bench_chain.err; echo rc=$?
B2D_RPN_CHAINS=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_nochain.json 2> gpurun_out/bench_nochain.err; echo rc=$?
