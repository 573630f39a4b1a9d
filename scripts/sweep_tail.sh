# Dev sweep: does an un-starved RPN-target tail (own high-priority stream + slim sampler CTAs) let the shortened chain pay?
run() { env "$@" python bench.py --steps 40 --warmup 3 --no-dropin --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('%-95s %.1f us/step   in_flight_2 %.1f' % ('$*', d['ms_per_step']*1e3, d['in_flight_2']['ms_per_step']*1e3))
"; }
run B2D_PDL=0
run B2D_RPN_TAIL_PRIO=-5 B2D_SAMPLE_THREADS=128
run B2D_RPN_TAIL_PRIO=-5 B2D_SAMPLE_THREADS=128 B2D_PDL=1
run B2D_RPN_TAIL_PRIO=-5 B2D_SAMPLE_THREADS=128 B2D_FUSE_TARGETS=1
run B2D_RPN_TAIL_PRIO=-5 B2D_SAMPLE_THREADS=128 B2D_FUSE_TARGETS=1 B2D_PDL=1
run B2D_RPN_TAIL_PRIO=-5 B2D_SAMPLE_THREADS=1024 B2D_FUSE_TARGETS=1 B2D_PDL=1
run B2D_RPN_TAIL_PRIO=-5 B2D_SAMPLE_THREADS=128 B2D_FUSE_TARGETS=1 B2D_PDL=1 B2D_RPN_PRIO=-3
