#!/bin/bash
# round-2 diagnostic pass: GPU tests, headline benches, then per-role wait shares (STCD_TRACE) of SNUNet at 64 pairs and SiamUnet_diff at 8
mkdir -p gpurun_out
tag=${1:-d}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$tag.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-also > gpurun_out/bench_${tag}_default.log 2>&1; echo "bench rc=$?"
STCD_XF_FAST_MIN=99 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-also > gpurun_out/bench_${tag}_default_noxf.log 2>&1
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-also > gpurun_out/bench_${tag}_default2.log 2>&1
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-also --workload siamunet_diff_256 > gpurun_out/bench_${tag}_c1_b8.log 2>&1
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-also --workload siamunet_diff_256_b64 > gpurun_out/bench_${tag}_c1_b64.log 2>&1
python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
for f in ("default", "default_noxf", "default2", "c1_b8", "c1_b64"):
    try:
        line = [l for l in open(f"gpurun_out/bench_{tag}_{f}.log") if l.startswith("{")][-1]
        d = json.loads(line)
        print(f, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 3), "mhz", d["clocks"]["sm_mhz"],
              "roof", d["roofline"]["kernel"], d["roofline"]["frac"])
        if f.startswith("default"):
            print("   " + " ".join(f"{n}:{ms*1e3:.0f}" for n, ms in d["per_op_ms"]))
    except Exception as e:
        print(f, "FAILED", e)
        import subprocess
        print(subprocess.run(["tail", "-5", f"gpurun_out/bench_{tag}_{f}.log"], capture_output=True, text=True).stdout)
PY
STCD_TRACE_NET=snunet python tools/trace_op.py 64 > gpurun_out/trace_${tag}_snunet64.log 2>&1
python tools/trace_op.py 8 > gpurun_out/trace_${tag}_siam8.log 2>&1
tail -4 gpurun_out/trace_${tag}_siam8.log
