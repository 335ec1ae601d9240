#!/bin/bash
# round-2 iteration check: GPU tests, then the headline benches (C2 + C3 ride-along, C1 b8 / b64)
mkdir -p gpurun_out
tag=${1:-x}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$tag.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${tag}_default.log 2>&1; echo "bench rc=$?"
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-also --workload siamunet_diff_256 > gpurun_out/bench_${tag}_c1_b8.log 2>&1
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-also --workload siamunet_diff_256_b64 > gpurun_out/bench_${tag}_c1_b64.log 2>&1
STCD_GRAPH=0 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-also --workload siamunet_diff_256 > gpurun_out/bench_${tag}_c1_b8_nograph.log 2>&1
python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
for f in ("default", "c1_b8", "c1_b64", "c1_b8_nograph"):
    try:
        line = [l for l in open(f"gpurun_out/bench_{tag}_{f}.log") if l.startswith("{")][-1]
        d = json.loads(line)
        print(f, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "e2e_u8", round(d["e2e_u8"]["value"]), "ms", round(d["ms_per_step"], 3),
              "roof", d["roofline"]["kernel"], d["roofline"]["frac"], "also", round(d.get("also", {}).get("value", 0)))
        if f in ("default",):
            print("   " + " ".join(f"{n}:{ms*1e3:.0f}" for n, ms in d["per_op_ms"]))
    except Exception as e:
        print(f, "FAILED", e)
        import subprocess
        print(subprocess.run(["tail", "-5", f"gpurun_out/bench_{tag}_{f}.log"], capture_output=True, text=True).stdout)
PY
