"""profiles/r2_parity.txt from the record the GPU tests write (tests/parity.py -> gpurun_out/parity_r2.jsonl)."""
import json
import os
import sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(root, "gpurun_out", "parity_r2.jsonl")
rows = [json.loads(l) for l in open(src) if l.strip()]
out = ["# Parity record of the GPU tests (tests/parity.py), B200, round 2, final kernels.  One line per case: logit spread of the fp32 oracle,",
       "# max / rms error, error relative to the spread, change-map agreement over ALL pixels and over decided pixels (|oracle margin| > tol).",
       "# tf32:* = precision path (split-bf16 operands), the rest = bf16 path.",
       f"{'case':66s} {'n_logits':>9s} {'std':>6s} {'max err':>9s} {'rms/std':>8s} {'max/std':>8s} {'all px':>8s} {'decided':>8s}"]
for r in rows:
    out.append(f"{r['case'][:66]:66s} {r['n_logits']:9d} {r['logit_std']:6.3f} {r['max_abs_err']:9.2e} {100 * r['rms_over_std']:7.3f}% "
               f"{100 * r['max_over_std']:7.2f}% {r['agree_all']:8.5f} {r['agree_decided']:8.5f}")
open(os.path.join(root, "profiles", "r2_parity.txt"), "w").write("\n".join(out) + "\n")
print(f"{len(rows)} cases")
