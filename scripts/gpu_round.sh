# Round measurement: GPU tests, bench (graph), eager launch list, ncu --set full of the hot kernels.
# usage: bash scripts/gpu_round.sh <tag>     then: python scripts/profile_summary.py <tag> <reps,comma-separated> gpurun_out/launches_<tag>.csv gpurun_out/bench_<tag>.json
tag=${1:-r1}
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo bench_rc=$?
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_${tag}_ref.json 2> gpurun_out/bench_${tag}_ref.err; echo ref_rc=$?
B="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu"
$B > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 3 --warmup 3 --no-graph --no-cpu > gpurun_out/ncu.log 2>&1; echo ncu_rc=$?
N="ncu --set full --clock-control none --import-source on"
$N -k regex:k_roi_align_win -s 3 -c 1 -o gpurun_out/prof_${tag}_roi $B > gpurun_out/ncu_a.log 2>&1; echo rc=$?
$N -k regex:'k_nms_sweep|k_nms_mask_sym|k_nms_cut|k_nms_scan' -s 10 -c 5 -o gpurun_out/prof_${tag}_nms $B > gpurun_out/ncu_b.log 2>&1; echo rc=$?
$N -k regex:'k_label_rows|k_colmax_rect|k_roi_targets_small|k_merge_rank|k_sample|k_hist|k_compact|k_select' -s 34 -c 17 -o gpurun_out/prof_${tag}_tgt $B > gpurun_out/ncu_c.log 2>&1; echo rc=$?
python scripts/timeline.py 1 > gpurun_out/timeline_$tag.txt 2>&1; echo tl_rc=$?
