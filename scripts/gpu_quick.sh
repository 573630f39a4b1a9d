# quick GPU check: full gpu test suite (all failures shown), bench line, graph timeline
# usage: bash scripts/gpu_quick.sh <tag> [pytest -k expression]
tag=${1:-q}
kexpr=${2:-}
if [ -n "$kexpr" ]; then python -m pytest tests -m gpu -q -k "$kexpr" 2>&1 | tail -40; else python -m pytest tests -m gpu -q 2>&1 | tail -40; fi
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo bench_rc=$?
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_$tag.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, "e2e", d["e2e"]["value"], "frac", round(r["frac"], 3), "path", round(r["path_frac"], 3), r["stage_ms"])
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/bench_$tag.err").read()[-3000:])
PY
python scripts/timeline.py 1 > gpurun_out/timeline_$tag.txt 2>&1; echo tl_rc=$?; grep -v Warning gpurun_out/timeline_$tag.txt | head -45
