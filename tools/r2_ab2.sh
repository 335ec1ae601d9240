#!/bin/bash
bash tools/r2_ab.sh "STCD_XF_MAX_CS=32" "STCD_FOLD_X_MAX_N=256" "STCD_XF_MAX_CS=64" "STCD_XF_MAX_CS=64 STCD_PDL=0" "STCD_XF_MAX_CS=64 STCD_GRAPH=0" | grep -v "layout="
for w in segcd_r34_1024_b16 changegnn_v1_256_b32 changeformer_v6_256_b32 siamunet_diff_256 siamunet_diff_256_b64 segcd_r50_1024_b16; do
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-also --workload $w > gpurun_out/r2d_$w.log 2>&1
  python - $w <<'PY'
import json, sys
w = sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/r2d_{w}.log") if l.startswith("{")][-1])
    print(w, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "u8", round(d["e2e_u8"]["value"]), "ms", round(d["ms_per_step"], 3), "MHz", d["clocks"]["sm_mhz"])
except Exception as e:
    print(w, "FAILED", e)
PY
done
