"""One forward of a harness net — the command ncu wraps.
usage: run_once.py [net=SiamUnet_diff] [pairs=8] [H=256] [chunk=pairs] [reps=1]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stcd_b200 import synth
from stcd_b200.networks import CLASSES

name = sys.argv[1] if len(sys.argv) > 1 else "SiamUnet_diff"
pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 8
H = int(sys.argv[3]) if len(sys.argv) > 3 else 256
chunk = int(sys.argv[4]) if len(sys.argv) > 4 else pairs
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 1
if name == "SegCD":
    net = CLASSES[name]("resnet34")
elif name == "BASE_Transformer":        # registry key base_transformer_pos_s4_dd8
    net = CLASSES[name](3, 2, with_pos="learned", resnet_stages_num=4, token_len=4, enc_depth=1, dec_depth=8)
elif name in ("ChangeGNNV2", "ChangeGNNV2_Compare", "VIG_V20_2"):
    net = CLASSES[name]()
elif name == "DSIFN":                   # registry key IFNet
    net = CLASSES[name]()
else:
    net = CLASSES[name](3, 2)
net = synth.prepare_(net.eval(), name).cuda()
net.chunk_pairs = chunk
x1, x2 = synth.image_pairs(pairs, H, H)
x1, x2 = x1.cuda(), x2.cuda()
if os.environ.get("STCD_PROFILE_RANGE"):     # ncu --profile-from-start off: plan creation (with its autotuning launches) and a warm-up
    y = net(x1, x2)                          # forward stay outside the profiled range; exactly one forward is captured
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
for _ in range(reps):
    y = net(x1, x2)
torch.cuda.synchronize()
if os.environ.get("STCD_PROFILE_RANGE"):
    torch.cuda.profiler.stop()
y = y[-1] if isinstance(y, (tuple, list)) else y
print("ok", float(y.abs().mean()))
if os.environ.get("STCD_DUMP_OPS"):          # op list in launch order, for tools/ncu_layers.py
    import json
    from stcd_b200 import lowering as L
    plan = net.plan_for(x1)
    ops = []
    for op in plan.prog.ops:
        kernels = 1
        if isinstance(op, L.EcamHeadSpec):
            kernels = 2
        elif isinstance(op, L.GraphConvSpec):
            kernels = 4 + (2 if op.r > 1 else 0)
        ops.append({"name": op.name, "type": type(op).__name__, "kernels": kernels,
                    "bytes_per_pair": int(L.op_bytes_per_pair(plan.prog, op)), "macs_per_pair": int(getattr(op, "macs_per_pair", 0)),
                    "n_tile": getattr(op, "n_tile", 0), "xf_cs": getattr(op, "xf_cs", 0), "fold_cs": getattr(op, "fold_cs", 0)})
    json.dump({"net": name, "pairs": pairs, "chunk": chunk, "h": H, "ops": ops}, open(os.environ["STCD_DUMP_OPS"], "w"))
