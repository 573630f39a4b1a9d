# Round measurement: GPU tests, bench (graph), eager launch list, ncu --set full of the hot kernels.
# usage: bash scripts/gpu_round.sh <tag>
tag=${1:-r1}
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo bench_rc=$?
python bench.py --steps 3 --warmup 3 --no-graph --no-cpu > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 3 --warmup 3 --no-graph --no-cpu > gpurun_out/ncu.log 2>&1; echo ncu_rc=$?
python bench.py --steps 2 --warmup 3 --no-graph --no-cpu > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_roi_align|k_nms_mask|k_nms_scan|k_assign_label|k_assign_colmax|k_select|k_merge|k_compact|k_hist|k_sample' -s 80 -c 16 -o gpurun_out/prof_$tag python bench.py --steps 2 --warmup 3 --no-graph --no-cpu > gpurun_out/ncu_full.log 2>&1; echo ncu_full_rc=$?
