"""Per-op lower bounds for a lowered Program: tensor-issue time (measured tcgen05 SS-mode floors)
and HBM time (each source read once + outputs written once), for B pairs on one B200."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stcd_b200 import siamunet, synth
from stcd_b200.lowering import ConvSpec

def cyc(n):   # measured cycles per tcgen05.mma M=128 K=16 (tools/ubench/mma_rate.cu)
    return 45 if n <= 64 else (64 if n <= 128 else 128)

def model(prog, B, clk=1.9e9, sms=148, hbm=6549e9):
    tot_m = tot_h = 0
    rows = []
    for op in prog.ops:
        if not isinstance(op, ConvSpec):
            continue
        n_img = B * (2 if op.pair else op.img_mult)
        tiles = ((op.hg + 15) // 16) * ((op.wg + 7) // 8) * n_img
        mmas = sum(ph.n_blocks for ph in op.phases) * (op.kc // 16) * (op.cout_pad // op.n_tile)
        t_m = tiles * mmas * cyc(op.n_tile) / sms / clk
        byt = 0
        seen = set()
        for ch in op.chunks:
            key = (ch.src, ch.stream)
            if key in seen: continue
            seen.add(key)
            t = prog.tensors[op.srcs[ch.src]]
            byt += t.h * t.w * t.c * 2 * B * (2 if op.pair else op.img_mult)
        ho, wo = op.hg * op.osy, op.wg * op.osx
        for o, f in ((op.out0, 1), (op.out_raw, 1), (op.out_pool, 0.25), (op.out_diff, 0.5 if op.pair else 1)):
            if o: byt += ho * wo * op.cout * 2 * n_img * f
        if op.out_ext >= 0: byt += ho * wo * op.cout * 4 * n_img
        t_h = byt / hbm
        rows.append((op.name, op.n_tile, op.kc, tiles, t_m * 1e6, t_h * 1e6, 2 * op.macs_per_pair * B / 1e9))
        tot_m += t_m; tot_h += t_h
    return rows, tot_m, tot_h

if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    net = synth.randomize_(siamunet.SiamUnet_diff(3, 2).eval(), gain=synth.GAINS["SiamUnet_diff"])
    rows, tm, th = model(net.lower(256, 256), B)
    for r in rows:
        print(f"{r[0]:9s} N={r[1]:3d} kc={r[2]:2d} tiles={r[3]:6d} mma={r[4]:7.1f}us hbm={r[5]:7.1f}us  gflop={r[6]:.2f}")
    s = sum(max(r[4], r[5]) for r in rows)
    print(f"sum mma {tm*1e6:.0f}us  sum hbm {th*1e6:.0f}us  sum max {s:.0f}us -> {B/s*1e6:.0f} pairs/s")
