#!/bin/bash
# Which role binds each conv op?  STCD_DBG bits: 1 no MMAs (the commits still arrive), 2 no epilogue stores / residual adds.  (A build with bit 4 = no accumulator loads showed no op moving: profiles/r2_role_sweep.txt.)
# Results are garbage by construction; only per_op_ms matters.
mkdir -p gpurun_out
./tools/ubench/mma_rate.bin > gpurun_out/mma_rate_r2.log 2>&1
for d in 0 1 2 3; do
  STCD_DBG=$d python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/dbg_c2_$d.log 2>&1
done
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-also --workload siamunet_diff_256 > gpurun_out/r2_c1_b8_base.log 2>&1
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-also --workload siamunet_diff_256_b64 > gpurun_out/r2_c1_b64_base.log 2>&1
python - <<'PY'
import json, glob
tabs = {}
for d in (0, 1, 2, 3):
    try:
        line = [l for l in open(f"gpurun_out/dbg_c2_{d}.log") if l.startswith("{")][-1]
        tabs[d] = dict((n, ms) for n, ms, *_ in json.loads(line)["per_op_ms"])
    except Exception as e:
        print("dbg", d, "failed", e)
names = list(tabs[0])
print(f"{'op':20s}" + "".join(f"{'dbg' + str(d):>9s}" for d in tabs))
for n in names:
    print(f"{n:20s}" + "".join(f"{tabs[d].get(n, float('nan')) * 1e3:9.1f}" for d in tabs))
print(f"{'sum':20s}" + "".join(f"{sum(tabs[d].values()) * 1e3:9.1f}" for d in tabs))
PY
