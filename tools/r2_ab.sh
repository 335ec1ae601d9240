#!/bin/bash
# A/B in ONE call (boxes differ by +-8 % in clocks): each variant = env settings, default C2 bench, per-op table side by side
mkdir -p gpurun_out
./tools/ubench/mma_n.bin > gpurun_out/mma_n_r2.log 2>&1
i=0
for v in "$@"; do
  env $v python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/ab_$i.log 2>&1
  i=$((i+1))
done
python - "$@" <<'PY'
import json, sys
vs = sys.argv[1:]
tabs = []
for i, v in enumerate(vs):
    try:
        d = json.loads([l for l in open(f"gpurun_out/ab_{i}.log") if l.startswith("{")][-1])
        tabs.append((v, d["value"], dict((n, ms) for n, ms in d["per_op_ms"]), d["clocks"]["sm_mhz"]))
    except Exception as e:
        print(v, "FAILED", e); tabs.append((v, 0, {}, 0))
for i, (v, val, _, mhz) in enumerate(tabs):
    print(f"[{i}] {v}: {val:.0f} pairs/s @ {mhz} MHz")
names = list(tabs[0][2])
print(f"{'op':18s}" + "".join(f"{'[' + str(i) + ']':>8s}" for i in range(len(tabs))))
for n in names:
    print(f"{n:18s}" + "".join(f"{t[2].get(n, float('nan')) * 1e3:8.0f}" for t in tabs))
print(f"{'sum':18s}" + "".join(f"{sum(t[2].values()) * 1e3:8.0f}" for t in tabs))
PY
cat gpurun_out/mma_n_r2.log
