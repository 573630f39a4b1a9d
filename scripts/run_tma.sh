timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for c in 1.5 1.0 1.25; do
B2D_NMS_CUT=$c timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_cut$c.json 2> gpurun_out/bench_cut$c.err; echo rc=$?
done
