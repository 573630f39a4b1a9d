# K6: tests + per-kernel times + end-to-end backward time
python -m pytest tests -m gpu -q -k "backward or bwd or reference_layout or cascade or extractor" 2>&1 | tail -5
python scripts/bench_bwd.py 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:bwd -c 12 --csv --log-file gpurun_out/bwd_launches.csv python scripts/bench_bwd.py > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/bwd_launches.csv")) if len(r)>5]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
for r in rows[-7:]: print(r[ki][:40], r[vi])
PY
