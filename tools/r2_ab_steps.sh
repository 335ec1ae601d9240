#!/bin/bash
# whole-step A/B in ONE call: variants (env settings) run round-robin REPS times; prints pairs/s and ms per step of every run
REPS=${REPS:-3}; STEPS=${STEPS:-30}; WORKLOAD=${WORKLOAD:-snunet_256_b64}
mkdir -p gpurun_out
for r in $(seq 1 $REPS); do
  i=0
  for v in "$@"; do
    env $v python bench.py --steps $STEPS --warmup 5 --no-cpu-baseline --no-also --workload $WORKLOAD > gpurun_out/abs_${i}_$r.log 2>&1
    python - "$v" gpurun_out/abs_${i}_$r.log <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[2]) if l.startswith("{")][-1])
    print(f"{sys.argv[1]:60s} {d['value']:8.0f} pairs/s  {d['ms_per_step']:7.3f} ms  e2e {d['e2e']['value']:8.0f}  @ {d['clocks']['sm_mhz']} MHz")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
    i=$((i+1))
  done
done
