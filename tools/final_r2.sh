#!/bin/bash
# Final round-2 evidence pass (ONE gpurun call): GPU tests, every workload's bench line, the reference arm, per-layer ncu tables of the
# headline configurations and the source-level capture of the dominant kernel.  Every ncu command runs only after the same command line
# exited 0 without ncu.
O=gpurun_out; mkdir -p $O
PARTS=${PARTS:-"tests bench prof trace"}
has() { [[ " $PARTS " == *" $1 "* ]]; }
if has tests; then
python -m pytest tests -m gpu -x -q > $O/pytest_fin.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_fin.log
fi
if has bench; then
python bench.py > $O/bench_fin_default.log 2>&1; echo "bench default rc=$?"
python bench.py --steps 300 --warmup 5 --no-cpu-baseline --no-also > $O/bench_fin_long.log 2>&1
bash tools/r2_multi.sh fin siamunet_diff_256 siamunet_diff_256_b64 segcd_r34_1024_b16 segcd_r50_1024_b16 changegnn_v1_256_b32 changeformer_v6_256_b32
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_fin_ref.log 2>&1; tail -1 $O/bench_fin_ref.log | cut -c1-300
./tools/ubench/mma_dyn.bin > $O/mma_dyn_r2.log 2>&1
fi
if has prof; then
bash tools/profile_r2.sh "c2_b64 SNUNet_ECAM 64 256 64" "c1_b8 SiamUnet_diff 8 256 8" "c1_b64 SiamUnet_diff 64 256 64" "c3_b4 SegCD 4 1024 4"
python tools/run_once.py SNUNet_ECAM 64 256 64 > $O/plain_src.log 2>&1 && {
  STCD_PROFILE_RANGE=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_ws -s 38 -c 1 -f -o $O/prof_fin_conv0_4_conv1 python tools/run_once.py SNUNet_ECAM 64 256 64 > $O/ncu_src_fin.log 2>&1
  python tools/ncu_src.py $O/prof_fin_conv0_4_conv1.ncu-rep 25 > $O/r2b_src_conv0_4_conv1.txt 2>&1
  head -2 $O/r2b_src_conv0_4_conv1.txt; tail -8 $O/r2b_src_conv0_4_conv1.txt
  rm -f $O/prof_fin_conv0_4_conv1.ncu-rep
}
fi
if has trace; then
# per-role wait tables: the TRACE build of the same sources (tools/trace_build.sh; built before the call, travels with the snapshot)
if [ -f stcd_b200/libstcd_b200_trace.so ]; then
  export STCD_LIB=stcd_b200/libstcd_b200_trace.so
  STCD_TRACE_NET=snunet python tools/trace_op.py 64 > $O/trace_fin_snunet64.log 2>&1
  STCD_TRACE_NET=segcd python tools/trace_op.py 4 > $O/trace_fin_segcd4.log 2>&1
  python tools/trace_op.py 8 > $O/trace_fin_siam8.log 2>&1
  python tools/trace_op.py 64 > $O/trace_fin_siam64.log 2>&1
  unset STCD_LIB
  for t in snunet64 segcd4 siam8 siam64; do python tools/trace_table.py $O/trace_fin_$t.log > $O/r2_roles_$t.txt; done
  tail -3 $O/r2_roles_siam8.txt
fi
fi
python - <<'PY'
import json
for f in ("default", "long"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/bench_fin_{f}.log") if l.startswith("{")][-1])
        print(f, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 3), "mhz", d["clocks"]["sm_mhz"], "roof", d["roofline"]["kernel"], d["roofline"]["frac"],
              "whole", d["roofline"].get("whole_step", {}).get("frac"), "also", round(d.get("also", {}).get("value", 0)), "cpu", d.get("cpu_baseline", {}).get("value"))
    except Exception as e:
        print(f, "FAILED", e)
PY
