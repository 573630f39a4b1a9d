# last check of a round: GPU tests, smoke, bench (both arms) exactly as the driver runs them
set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 | cut -c1-200
python bench.py --gpus 1 --steps 20 --warmup 3 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k: d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['path_frac'], d['cpu_baseline']['value'])"
