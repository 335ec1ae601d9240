"""ChangeGNNV1 (ViG pyramid encoder + multi-scale difference decoder, config C4) on the GPU against the oracle, the
emulator and the golden fixture generated from the reference's own ChangeVIG.py (with the restated gcn_lib: the
Grapher's parity is unpinned, see oracle/gcn.py)."""
import os

import numpy as np
import pytest
import torch

import parity
from oracle import emulate, nets
from stcd_b200 import changevig, synth

pytestmark = pytest.mark.gpu
BF16_TOL = 2e-2


def _net():
    return synth.prepare_(changevig.ChangeGNNV1().eval(), "ChangeGNNV1")


def _check(ys, refs, tol=BF16_TOL):
    assert isinstance(ys, list) and len(ys) == 5
    for y, ref in zip(ys, refs):
        assert y.shape == ref.shape and y.dtype == torch.float32
        assert (y.cpu() - ref).abs().max().item() < tol
    y, ref = ys[-1].cpu(), refs[-1]
    margin = (ref[:, 1] - ref[:, 0]).abs()
    agree = (y[:, 1] > y[:, 0]) == (ref[:, 1] > ref[:, 0])
    assert agree[margin > BF16_TOL].float().mean().item() >= 0.999
    assert agree.float().mean().item() >= 0.97
    # recorded for the parity table (tests/parity.py); relative bounds for a deep bf16 net (see tests/test_gpu_segcd.py)
    r = parity.report("%s:full-res logits" % "changegnn_v1", y, ref, "argmax")
    assert r["rms_over_std"] <= 0.04 and r["max_over_std"] <= 0.25, r


def test_forward_matches_oracle_and_emulator():
    net = _net()
    x1, x2 = synth.image_pairs(3, 256, 256)
    with torch.no_grad():
        ref = nets.changegnn_forward(net.state_dict(), x1, x2)
    emu = emulate.run_program(net.lower(256, 256), x1[:1], x2[:1], chunk=1)
    net = net.cuda()
    net.chunk_pairs = 2                      # 3 pairs -> one full chunk + a ragged one
    ys = net(x1.cuda(), x2.cuda())
    _check(ys, ref)
    for y, e in zip(ys, emu):
        assert (y[:1].cpu() - e).abs().max().item() < 1.5e-2, "kernel vs emulator (same rounding points)"


def test_forward_matches_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "changegnn_v1.npz"))
    assert float(g["gain"]) == synth.GAINS["ChangeGNNV1"] and int(g["n_out"]) == 5
    net = _net().cuda()
    x1, x2 = synth.image_pairs(int(g["batch"]), int(g["h"]), int(g["w"]), seed=int(g["data_seed"]))
    ys = net(x1.cuda(), x2.cuda())
    _check(ys, [torch.from_numpy(g[f"out{i}"]) for i in range(5)])


def test_layerwise_against_emulator():
    net = _net()
    x1, x2 = synth.image_pairs(1, 256, 256, seed=9)
    keep = {}
    emulate.run_program(net.lower(256, 256), x1, x2, chunk=1, keep=keep)
    net = net.cuda()
    net.chunk_pairs = 1
    net(x1.cuda(), x2.cuda())
    torch.cuda.synchronize()
    plan = net.plan_for(x1.cuda())
    worst = {}
    for name in plan.prog.tensors:
        got = plan.read_tensor(name)
        worst[name] = ((got - keep[name]).abs().mean() / (keep[name].abs().mean() + 1e-6)).item()
    # the kNN choice is discrete: where the (bf16-drifted) inputs of a Grapher differ by an ulp, a near-equidistant
    # neighbour swaps and the max-relative tensor m (and the grouped conv gz right behind it) moves by more than the
    # usual drift; everything else stays within a few percent in the mean
    def limit(name):
        return 0.2 if name.endswith((".m", ".gz")) else 4e-2
    bad = {k: v for k, v in worst.items() if v > limit(k)}
    assert not bad, bad


def test_properties_and_conventions():
    net = _net().cuda()
    x1, x2 = synth.image_pairs(4, 256, 256, seed=3)
    a, b = x1.cuda(), x2.cuda()
    ys = net(a, b)
    ys2 = net(a, b)
    assert all(torch.equal(u, v) for u, v in zip(ys, ys2)), "forward must be deterministic"
    perm = torch.tensor([2, 0, 3, 1])
    yp = net(a[perm], b[perm])
    assert all(torch.equal(u, v[perm.cuda()]) for u, v in zip(yp, ys)), "pairs are independent"
    assert [tuple(t.shape[1:]) for t in ys] == [(2, 8, 8), (2, 16, 16), (2, 32, 32), (2, 64, 64), (2, 256, 256)]
    from types import SimpleNamespace
    from stcd_b200.networks import define_G
    g = define_G(SimpleNamespace(net_G="ChangeGNNV1", n_class=2, embed_dim=256, img_size=256), gpu_ids=[0])
    assert isinstance(g, changevig.ChangeGNNV1)
    with pytest.raises(ValueError):
        net(torch.zeros(1, 3, 128, 128).cuda(), torch.zeros(1, 3, 128, 128).cuda())
