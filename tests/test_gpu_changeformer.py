"""ChangeFormerV6 (MiT transformer encoder + multi-scale difference decoder, config C5) on the GPU against the oracle, the
emulator and the golden fixture generated from the unmodified reference ChangeFormer.py (timm's DropPath shimmed:
identity in eval mode)."""
import os

import numpy as np
import pytest
import torch

import parity
from oracle import emulate, nets
from stcd_b200 import changeformer, synth

pytestmark = pytest.mark.gpu
BF16_TOL = 2e-2


def _net():
    return synth.prepare_(changeformer.ChangeFormerV6().eval(), "ChangeFormerV6")


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
    r = parity.report("%s:full-res logits" % "changeformer_v6", y, ref, "argmax")
    assert r["rms_over_std"] <= 0.04 and r["max_over_std"] <= 0.25, r


def test_forward_matches_oracle_and_emulator():
    net = _net()
    x1, x2 = synth.image_pairs(3, 256, 256)
    with torch.no_grad():
        ref = nets.changeformer_forward(net.state_dict(), x1, x2)
    emu = emulate.run_program(net.lower(256, 256), x1[:1], x2[:1], chunk=1)
    net = net.cuda()
    net.chunk_pairs = 2                      # 3 pairs -> one full chunk + a ragged one
    ys = net(x1.cuda(), x2.cuda())
    _check(ys, ref)
    for y, e in zip(ys, emu):
        assert (y[:1].cpu() - e).abs().max().item() < 1.5e-2, "kernel vs emulator (same rounding points)"


def test_forward_matches_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "changeformer_v6.npz"))
    assert float(g["gain"]) == synth.GAINS["ChangeFormerV6"] and int(g["n_out"]) == 5
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
    def limit(name):
        return 4e-2
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
    g = define_G(SimpleNamespace(net_G="ChangeFormerV6", n_class=2, embed_dim=256, img_size=256), gpu_ids=[0])
    assert isinstance(g, changeformer.ChangeFormerV6)
    with pytest.raises(ValueError):
        net(torch.zeros(1, 3, 128, 128).cuda(), torch.zeros(1, 3, 128, 128).cuda())


@pytest.mark.parametrize("version", ["1", "2", "3"])
def test_changeformer_v1_v2_match_oracle_and_golden(version, golden_dir):
    """ChangeFormerV1 / V2 (Tenc encoder, |fx1 - fx2|, convprojection_base / TDec): ONE tensor out, like upstream."""
    import os

    import numpy as np
    from types import SimpleNamespace

    from stcd_b200 import changeformer as cf, networks
    cls = f"ChangeFormerV{version}"
    net = synth.prepare_(getattr(cf, cls)().eval(), cls)
    x1, x2 = synth.image_pairs(3, 256, 256)
    with torch.no_grad():
        ref = getattr(nets, f"changeformer_v{version}_forward")(net.state_dict(), x1, x2)
    net = net.cuda()
    net.chunk_pairs = 2                      # 3 pairs -> one full chunk + a ragged one
    y = net(x1.cuda(), x2.cuda())
    assert isinstance(y, torch.Tensor) and y.shape == ref.shape and y.dtype == torch.float32
    y = y.cpu()

    def decided(a, b):
        margin = (b[:, 1] - b[:, 0]).abs()
        return ((a[:, 1] > a[:, 0]) == (b[:, 1] > b[:, 0]))[margin > 2e-2].float().mean().item()

    assert (y - ref).abs().max().item() < 2e-2 and decided(y, ref) >= 0.999
    assert 0.02 < (ref[:, 1] > ref[:, 0]).float().mean().item() < 0.98
    g = np.load(os.path.join(golden_dir, f"changeformer_v{version}.npz"))
    x1, x2 = synth.image_pairs(int(g["batch"]), int(g["h"]), int(g["w"]), seed=int(g["data_seed"]))
    yg = net(x1.cuda(), x2.cuda()).cpu()
    gref = torch.from_numpy(g["out0"])
    assert (yg - gref).abs().max().item() < 2e-2 and decided(yg, gref) >= 0.999
    n2 = networks.define_G(SimpleNamespace(net_G=cls, n_class=2, embed_dim=256, img_size=256), gpu_ids=[0])
    assert isinstance(n2, getattr(cf, cls)) and next(n2.parameters()).is_cuda
