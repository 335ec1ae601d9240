"""IFNet (DSIFN) on the GPU against the oracle, the emulator and the golden fixture generated from the unmodified
reference (models/DSIFN.py): logits within 2e-2 absolute (bf16 path), change maps (sigmoid(out) > 0.5) agreeing on
>= 99.9 % of decided pixels."""
import os

import numpy as np
import pytest
import torch

from oracle import emulate, nets
from stcd_b200 import dsifn, networks, synth

pytestmark = pytest.mark.gpu
BF16_TOL = 2e-2


def _net():
    return synth.prepare_(networks.CLASSES["DSIFN"]().eval(), "DSIFN")


def _decided_agreement(y, ref):
    return ((y > 0) == (ref > 0))[ref.abs() > BF16_TOL].float().mean().item()


def test_forward_matches_oracle_and_emulator():
    net = _net()
    x1, x2 = synth.image_pairs(3, 64, 96)
    with torch.no_grad():
        ref = nets.dsifn_forward(net.state_dict(), x1, x2)
    emu = emulate.run_program(net.lower(64, 96), x1[:2], x2[:2], chunk=2)[0]
    net = net.cuda()
    net.chunk_pairs = 2                      # 3 pairs -> one full chunk + a ragged one
    y = net(x1.cuda(), x2.cuda())
    assert isinstance(y, torch.Tensor) and y.shape == ref.shape == (3, 1, 64, 96) and y.dtype == torch.float32
    y = y.cpu()
    assert (y[:2] - emu).abs().max().item() < 1.5e-2, "kernel vs emulator (same rounding points; bf16 flips cascade)"
    assert (y - ref).abs().max().item() < BF16_TOL, "kernel vs fp32 oracle"
    assert _decided_agreement(y, ref) >= 0.999
    assert 0.02 < (ref > 0).float().mean().item() < 0.98


def test_intermediate_tensors_match_emulator():
    """Every plan tensor of the first chunk against the emulator's: localises a wrong op (gates, virtual concat, pools)."""
    net = _net()
    x1, x2 = synth.image_pairs(2, 64, 64)
    keep = {}
    emulate.run_program(net.lower(64, 64), x1, x2, chunk=2, keep=keep)
    net = net.cuda()
    net.chunk_pairs = 2
    net(x1.cuda(), x2.cuda())
    torch.cuda.synchronize()
    plan = net.plan_for(x1.cuda())
    bad = []
    for name in plan.prog.tensors:
        got, want = plan.read_tensor(name), keep[name]
        err = ((got - want).abs().mean() / (want.abs().mean() + 1e-3)).item()
        if err > 0.03:
            bad.append((name, err))
    assert not bad, bad[:8]


def test_forward_matches_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "ifnet.npz"))
    net = _net().cuda()
    x1, x2 = synth.image_pairs(int(g["batch"]), int(g["h"]), int(g["w"]), seed=int(g["data_seed"]))
    y = net(x1.cuda(), x2.cuda()).cpu()
    ref = torch.from_numpy(g["out0"])
    assert (y - ref).abs().max().item() < BF16_TOL
    assert _decided_agreement(y, ref) >= 0.999


def test_256_shape_and_properties():
    net = _net()
    x1, x2 = synth.image_pairs(5, 256, 256)
    with torch.no_grad():
        ref = nets.dsifn_forward(net.state_dict(), x1[:1], x2[:1])
    net = net.cuda()
    net.chunk_pairs = 4
    y = net(x1.cuda(), x2.cuda())
    assert (y[:1].cpu() - ref).abs().max().item() < BF16_TOL
    assert torch.equal(y, net(x1.cuda(), x2.cuda())), "forward must be deterministic"
    perm = torch.tensor([3, 1, 4, 0, 2])
    yp = net(x1[perm].cuda(), x2[perm].cuda())
    assert torch.equal(yp, y[perm.cuda()]), "pairs are independent: permuting the batch permutes the logits"
    plan = net.plan_for(x1.cuda())
    outs = plan.forward_host(x1.pin_memory(), x2.pin_memory())
    assert torch.equal(outs[0], y.cpu()), "host-buffer path must equal the device path bit for bit"


def test_define_G_returns_ifnet():
    from types import SimpleNamespace
    net = networks.define_G(SimpleNamespace(net_G="IFNet", n_class=2), gpu_ids=[0])
    assert isinstance(net, dsifn.DSIFN) and net.t1_base is net.t2_base and next(net.parameters()).is_cuda
