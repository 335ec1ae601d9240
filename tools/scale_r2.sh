#!/bin/bash
# 1 -> 8 GPU scaling of bench.py as the driver runs it (gpurun --gpus 8 -- bash tools/scale_r2.sh)
mkdir -p gpurun_out
for n in 1 2 4 8; do
  if [ $n = 1 ]; then
    python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/scale_r2_n$n.log 2>&1
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29510 + n)) bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/scale_r2_n$n.log 2>&1
  fi
done
python - <<'PY'
import json
base = {}
for n in (1, 2, 4, 8):
    try:
        d = json.loads([l for l in open(f"gpurun_out/scale_r2_n{n}.log") if l.startswith("{")][-1])
    except Exception as e:
        print(n, "FAILED", e); continue
    a = d.get("also", {})
    row = {"C2 value": d["value"], "C2 e2e": d["e2e"]["value"], "C2 e2e_u8": d["e2e_u8"]["value"],
           "C3 value": a.get("value", 0), "C3 e2e(u8)": a.get("e2e", {}).get("value", 0), "C3 e2e_f32": a.get("e2e_f32", {}).get("value", 0)}
    if n == 1:
        base = row
    print(f"N={n} " + "  ".join(f"{k} {v:8.0f} (eff {v / (n * base[k]):.3f})" if base.get(k) else f"{k} {v:8.0f}" for k, v in row.items()), d.get("host"))
PY
