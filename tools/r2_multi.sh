#!/bin/bash
# one line per workload: value / ms per step / e2e (bench.py, 20 steps)
mkdir -p gpurun_out
tag=${1:-m}; shift
for w in "$@"; do
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-also --workload $w > gpurun_out/bench_${tag}_$w.log 2>&1
  python - gpurun_out/bench_${tag}_$w.log $w <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    print(f"{sys.argv[2]:28s} value {d['value']:9.1f}  ms {d['ms_per_step']:8.3f}  e2e {d['e2e']['value']:9.1f}  e2e_u8 {d.get('e2e_u8', {}).get('value', 0):9.1f}  @ {d['clocks']['sm_mhz']} MHz  roof {d['roofline']['kernel']} {d['roofline']['frac']}")
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
done
