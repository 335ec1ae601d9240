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
    # usage: floor_model.py [net] [B] [H] [bench log with per_op_ms]
    import json
    from stcd_b200.networks import CLASSES
    name = sys.argv[1] if len(sys.argv) > 1 else "SiamUnet_diff"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    H = int(sys.argv[3]) if len(sys.argv) > 3 else 256
    net = CLASSES[name]("resnet34") if name == "SegCD" else CLASSES[name](3, 2)
    prog = net.eval().lower(H, H)
    meas = {}
    if len(sys.argv) > 4:
        line = [l for l in open(sys.argv[4]) if l.startswith("{")][-1]
        meas = {n: ms for n, ms, *_ in json.loads(line)["per_op_ms"]}
    rows, tm, th = model(prog, B)
    print(f"{'op':34s} {'N':>4s} {'kc':>3s} {'tiles':>7s} {'mma us':>8s} {'hbm us':>8s} {'GFLOP':>8s} {'meas us':>8s} {'TF/s':>7s} {'x floor':>7s}")
    tot = 0
    for r in rows:
        ms = meas.get(r[0])
        fl = max(r[4], r[5])
        extra = f" {ms * 1e3:8.1f} {r[6] / ms:7.1f} {ms * 1e3 / fl:7.2f}" if ms else ""
        tot += ms or 0
        print(f"{r[0]:34s} {r[1]:4d} {r[2]:3d} {r[3]:7d} {r[4]:8.1f} {r[5]:8.1f} {r[6]:8.1f}{extra}")
    print(f"sum of floors: mma {tm * 1e6:.0f} us, hbm {th * 1e6:.0f} us, max-per-op {sum(max(r[4], r[5]) for r in rows):.0f} us; measured convs {tot * 1e3:.0f} us")
