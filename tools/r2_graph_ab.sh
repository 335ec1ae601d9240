#!/bin/bash
mkdir -p gpurun_out
for w in changegnn_v1_256_b32 segcd_r34_1024_b16 snunet_256_b64 siamunet_diff_256; do
  for g in 1 0; do
    STCD_GRAPH=$g python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-also --workload $w > gpurun_out/gab_${w}_$g.log 2>&1
    python - $w $g <<'PY'
import json, sys
w, g = sys.argv[1:3]
try:
    d = json.loads([l for l in open(f"gpurun_out/gab_{w}_{g}.log") if l.startswith("{")][-1])
    print(w, "graph", g, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "u8", round(d["e2e_u8"]["value"]), "ms", round(d["ms_per_step"], 3), "MHz", d["clocks"]["sm_mhz"], "sum-of-ops ms", round(sum(ms for _, ms in d["per_op_ms"]), 3))
except Exception as e:
    print(w, g, "FAILED", e)
PY
  done
done
