# Dev sweep: step time of the config-2 graph under the chain-shortening knobs (PDL, RoI targets as the proposal kernel's
# tail) and the priority of the RPN-target stream.  usage: bash scripts/sweep_step_knobs.sh
for prio in 0 -3; do for pdl in 0 1; do for fuse in 0 1; do
  B2D_RPN_PRIO=$prio B2D_PDL=$pdl B2D_FUSE_TARGETS=$fuse python bench.py --steps 40 --warmup 3 --no-dropin --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('rpn_prio %3s  pdl %s  fuse_targets %s   %.1f us/step   in_flight_2 %.1f' % ('$prio', '$pdl', '$fuse', d['ms_per_step']*1e3, d['in_flight_2']['ms_per_step']*1e3))
"
done; done; done
