"""Layer-by-layer comparison of the GPU plan's activation tensors with the CPU emulator (debug aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import emulate
from stcd_b200 import siamunet, synth
from stcd_b200.plan import Plan

fusion = sys.argv[1] if len(sys.argv) > 1 else "diff"
H, W, B, chunk = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
cls = {"diff": siamunet.SiamUnet_diff, "conc": siamunet.SiamUnet_conc}[fusion]
net = synth.randomize_(cls(3, 2).eval(), gain=synth.GAINS["SiamUnet_diff"])
x1, x2 = synth.image_pairs(B, H, W)
prog = net.lower(H, W)
keep = {}
emu = emulate.run_program(prog, x1, x2, chunk=chunk, keep=keep)[0]
plan = Plan(prog, chunk)
y = plan.forward(x1.cuda(), x2.cuda())[0].cpu()
print("logits max diff", (y - emu).abs().max().item())
for name in prog.tensors:
    g = plan.read_tensor(name)
    e = keep[name]
    d = (g - e).abs()
    bad = (d > 0.0079 * e.abs() + 2e-3)
    print(f"{name:8s} {tuple(g.shape)} max {d.max().item():.4g} bad {bad.float().mean().item():.4%}",
          ("first bad idx " + str(bad.nonzero()[0].tolist())) if bad.any() else "")
