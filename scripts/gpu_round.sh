set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1d.json 2> gpurun_out/bench_r1d.err; echo bench_rc=$?
python bench.py --steps 3 --warmup 3 --no-graph --no-cpu > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1d.csv python bench.py --steps 3 --warmup 3 --no-graph --no-cpu > gpurun_out/ncu.log 2>&1; echo ncu_rc=$?
python bench.py --steps 2 --warmup 3 --no-graph --no-cpu > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_roi_align_nhwc|k_nms_mask|k_nms_scan|k_assign_label|k_assign_colmax|k_select|k_merge|k_compact|k_hist' -s 60 -c 14 -o gpurun_out/prof_r1d python bench.py --steps 2 --warmup 3 --no-graph --no-cpu > gpurun_out/ncu_full.log 2>&1; echo ncu_full_rc=$?
