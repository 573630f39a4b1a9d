# sparse host->device fetch: tests + bench line
python -m pytest tests -m gpu -q -x -k "mark_and_fetch or sparse_fetch or train_path" 2>&1 | tail -15
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r2m.json 2> gpurun_out/bench_r2m.err; echo bench_rc=$?
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_r2m.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")})
    print("e2e", d["e2e"]); print("e2e_nchw", d["e2e_nchw"]); print("feats", d["e2e_with_roi_feats"])
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/bench_r2m.err").read()[-3000:])
PY
