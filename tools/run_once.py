"""One SiamUnet_diff forward (chunk pairs = argv[1], default 8) — the command ncu wraps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stcd_b200 import siamunet, synth

chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
net = synth.randomize_(siamunet.SiamUnet_diff(3, 2).eval(), gain=synth.GAINS["SiamUnet_diff"]).cuda()
net.chunk_pairs = chunk
x1, x2 = synth.image_pairs(chunk, 256, 256)
x1, x2 = x1.cuda(), x2.cuda()
for _ in range(reps):
    y = net(x1, x2)
torch.cuda.synchronize()
print("ok", float(y.abs().mean()))
