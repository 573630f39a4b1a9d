# Round-2 measurement pass: GPU tests, bench (graph), eager launch list, ncu --set full of the hot kernels, graph timeline.
# usage: bash scripts/gpu_round2.sh <tag>   then: python scripts/profile_summary.py <tag> <reps,comma-separated> gpurun_out/launches_<tag>.csv gpurun_out/bench_<tag>.json
tag=${1:-r2}
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo bench_rc=$?
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_${tag}_ref.json 2> gpurun_out/bench_${tag}_ref.err; echo ref_rc=$?
B="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu --no-dropin"
$B > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 3 --warmup 3 --no-graph --no-cpu --no-dropin > gpurun_out/ncu.log 2>&1; echo ncu_rc=$?
N="ncu --set full --clock-control none --import-source on"
$N -k regex:k_roi_align_win -s 3 -c 1 -o gpurun_out/prof_${tag}_roi $B > gpurun_out/ncu_a.log 2>&1; echo rc=$?
$N -k regex:'k_rpn_front|k_rpn_back' -s 6 -c 3 -o gpurun_out/prof_${tag}_rpn $B > gpurun_out/ncu_b.log 2>&1; echo rc=$?
$N -k regex:'k_label_rows|k_colmax_rect|k_roi_targets_small|k_sample|k_encode_targets|k_gather_head' -s 12 -c 6 -o gpurun_out/prof_${tag}_tgt $B > gpurun_out/ncu_c.log 2>&1; echo rc=$?
$N -k regex:'k_fetch_cells|k_roi_mark' -c 2 -o gpurun_out/prof_${tag}_fetch $B > gpurun_out/ncu_d.log 2>&1; echo rc=$?
python scripts/timeline.py 1 > gpurun_out/timeline_$tag.txt 2>&1; echo tl_rc=$?
python scripts/bench_rpn.py "B2D_DBG=10" > gpurun_out/rpn_phases_$tag.txt 2>&1; echo ph_rc=$?
python scripts/bench_roi_order.py > gpurun_out/roi_order_$tag.txt 2>&1; echo ord_rc=$?
python scripts/bench_dropin.py > gpurun_out/dropin_phases_$tag.txt 2>&1; echo dp_rc=$?
python scripts/bench_cascade.py > gpurun_out/cascade_$tag.txt 2>&1; echo cas_rc=$?
