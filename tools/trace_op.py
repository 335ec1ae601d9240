"""Per-CTA clock stamps of the conv ops of one chunk (STCD_TRACE=1): SiamUnet_diff, or STCD_TRACE_NET=snunet / segcd.
The per-role wait counters (issuer / epilogue / producer) are compiled into the TRACE build only:
    bash tools/trace_build.sh && STCD_LIB=stcd_b200/libstcd_b200_trace.so python tools/trace_op.py 64
(with the product library the stamps are there and the wait shares read 0)."""
import os, sys
os.environ["STCD_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
import torch
from stcd_b200 import siamunet, synth, _lib
from stcd_b200.plan import Plan

H = W = 256
chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 8
from stcd_b200 import snunet
which = os.environ.get("STCD_TRACE_NET", "siam")
if which == "segcd":
    from stcd_b200 import smp
    H = W = 1024
    net = synth.prepare_(smp.SegCD("resnet34", classes=1).eval(), "SegCD")
else:
    net = synth.prepare_(snunet.SNUNet_ECAM(3, 2).eval(), "SNUNet_ECAM") if which == "snunet" else synth.randomize_(siamunet.SiamUnet_diff(3, 2).eval(), gain=synth.GAINS["SiamUnet_diff"])
prog = net.lower(H, W)
plan = Plan(prog, chunk)
x1, x2 = synth.image_pairs(chunk, H, W)
x1, x2 = x1.cuda(), x2.cuda()
for _ in range(3):
    plan.forward(x1, x2)
torch.cuda.synchronize()
lib = _lib.lib()
names = ["gt", "setup", "mma:a0", "mma:wait_w", "mma:t0", "epi:t0beg", "epi:t0end", "exit", "prod:tma0", "mma:last", "epi:last", "tiles", "mma:wait_a", "mma:wait_acc", "epi:wait_acc", "prod:wait_empty"]
for i, op in enumerate(prog.ops):
    info = (C.c_int32 * 10)()
    buf = np.zeros(1 << 16, dtype=np.int64)
    n = lib.stcd_plan_read_trace(plan._h, i, buf.ctypes.data, buf.size, info)
    if n <= 0:
        continue
    t = buf[:n].reshape(-1, 16)
    gt = t[:, 0]
    span_us = (gt.max() - gt.min()) / 1e3
    end_ns = (gt + t[:, 7] / 1.965).max()
    if "t_origin" not in globals():
        t_origin = gt.min()
        prev_end = 0
    first_mma_ns = (gt + t[:, 2] / 1.965).min()
    print(f"     timeline: start {(gt.min() - t_origin) / 1e3:7.1f} us  first-mma {(first_mma_ns - t_origin) / 1e3:7.1f}  end {(end_ns - t_origin) / 1e3:7.1f}  "
          f"(start - prev end {(gt.min() - prev_end - t_origin) / 1e3 if prev_end else 0:6.1f}, first-mma - prev end {(first_mma_ns - prev_end - t_origin) / 1e3 if prev_end else 0:6.1f}, busy {(end_ns - max(first_mma_ns, prev_end + t_origin)) / 1e3:6.1f})")
    prev_end = end_ns - t_origin
    med = np.median(t, axis=0)
    print(f"{op.name:8s} grid=({info[0]},{info[1]}) smem={info[2]} aS={info[3]} wS={info[4]} res={info[5]} tmem={info[6]} tiles={info[7]} aB={info[8]} wB={info[9]} start-span={span_us:.1f}us")
    print("     median cycles: " + " ".join(f"{names[k]}={int(med[k])}" for k in (1, 8, 2, 4, 5, 6, 9, 10, 7, 11)))
    run = max(1.0, med[7] - med[2])          # first operand -> exit
    print("     share of (first operand -> exit) spent waiting: " + " ".join(f"{names[k]}={med[k] / run:.2f}" for k in (12, 3, 13, 14, 15)))
