"""The precision path (``net.precision = "tf32"``: split-bf16 operands, stcd_b200/lowering.py Program.precision) on the GPU:
logits within 1e-3 ABSOLUTE of the fp32 oracle -- the north star's tolerance for the tf32 path -- at a logit spread >= 0.25,
on configs C1 (SiamUnet_diff) and C2 (SNUNet-CD), through the same C-ABI and kernels as the bf16 path."""
import os

import numpy as np
import pytest
import torch

import parity
from oracle import emulate, nets
from stcd_b200 import siamunet, snunet, synth

pytestmark = pytest.mark.gpu
TF32_TOL = 1e-3


def _check(name, y, ref):
    r = parity.check(name, y, ref, "argmax", abs_tol=TF32_TOL, rms_rel=1e-3, max_rel=4e-3, all_px=0.9995, decided=1.0, min_std=0.25)
    assert r["max_abs_err"] < TF32_TOL
    return r


@pytest.mark.parametrize("fusion", ["diff", "conc", "sub", "cross", "ef"])
def test_siamunet_family_within_1e3(fusion):
    cls = {"diff": siamunet.SiamUnet_diff, "conc": siamunet.SiamUnet_conc, "sub": siamunet.SiamUnet_sub,
           "cross": siamunet.SiamUnet_cross_conc, "ef": siamunet.Unet}[fusion]
    net = synth.randomize_(cls(3, 2).eval(), gain=0.8)          # every variant at a logit spread >= 0.25
    net.precision = "tf32"
    x1, x2 = synth.image_pairs(5, 64, 96)
    with torch.no_grad():
        ref = nets.siamunet_forward(net.state_dict(), x1, x2, fusion)
    emu = emulate.run_program(net.lower(64, 96), x1, x2, chunk=4)[0]
    net = net.cuda()
    net.chunk_pairs = 4                      # 5 pairs -> one full chunk + a ragged one
    y = net(x1.cuda(), x2.cuda())
    y = (y[0] if isinstance(y, list) else y).cpu()
    _check(f"tf32:siamunet_{fusion}:5x64x96", y, ref)
    assert (y - emu).abs().max().item() < 2e-4, "kernel vs emulator (same split points, fp32 accumulation order differs)"


def test_snunet_within_1e3():
    net = synth.prepare_(snunet.SNUNet_ECAM(3, 2).eval(), "SNUNet_ECAM")
    net.precision = "tf32"
    x1, x2 = synth.image_pairs(3, 64, 96)
    with torch.no_grad():
        ref = nets.snunet_forward(net.state_dict(), x1, x2)
    emu = emulate.run_program(net.lower(64, 96), x1, x2, chunk=2)[0]
    net = net.cuda()
    net.chunk_pairs = 2
    y = net(x1.cuda(), x2.cuda()).cpu()
    _check("tf32:snunet:3x64x96", y, ref)
    assert (y - emu).abs().max().item() < 2e-4


def test_config_c1_and_c2_shapes_within_1e3():
    """Configs C1 (SiamUnet_diff, 8 pairs of 256x256) and C2's shape (SNUNet, 256x256) on the precision path, and the two paths
    of ONE module side by side: the bf16 plan and the tf32 plan are cached separately and switching back and forth is exact."""
    net = synth.prepare_(siamunet.SiamUnet_diff(3, 2).eval(), "SiamUnet_diff")
    x1, x2 = synth.image_pairs(8, 256, 256)
    with torch.no_grad():
        ref = nets.siamunet_forward(net.state_dict(), x1[:2], x2[:2], "diff")
    net = net.cuda()
    y_bf16 = net(x1.cuda(), x2.cuda())
    net.precision = "tf32"
    y = net(x1.cuda(), x2.cuda())
    _check("tf32:siamunet_diff:C1 8x256x256 (pairs 0-1 vs oracle)", y[:2], ref)
    assert torch.equal(y, net(x1.cuda(), x2.cuda())), "deterministic"
    r16 = parity.report("bf16:siamunet_diff:C1 (same pairs, for the comparison)", y_bf16[:2], ref, "argmax")
    assert r16["max_abs_err"] > 10 * (y[:2].cpu() - ref).abs().max().item(), "the precision path must be far tighter than bf16"
    net.precision = "bf16"
    assert torch.equal(net(x1.cuda(), x2.cuda()), y_bf16)

    sn = synth.prepare_(snunet.SNUNet_ECAM(3, 2).eval(), "SNUNet_ECAM")
    a, b = synth.image_pairs(4, 256, 256)
    with torch.no_grad():
        ref2 = nets.snunet_forward(sn.state_dict(), a[:2], b[:2])
    sn = sn.cuda()
    sn.precision = "tf32"
    sn.chunk_pairs = 4
    y2 = sn(a.cuda(), b.cuda())
    _check("tf32:snunet:C2 shape 4x256x256 (pairs 0-1 vs oracle)", y2[:2], ref2)


def test_define_G_precision_switch():
    from types import SimpleNamespace
    from stcd_b200 import networks
    net = networks.define_G(SimpleNamespace(net_G="SNUNet", n_class=2, precision="tf32"), gpu_ids=[0])
    assert net.precision == "tf32" and net.plan_precision == "split"
    x1, x2 = synth.image_pairs(1, 32, 32)
    y = net.eval()(x1.cuda(), x2.cuda())
    assert y.shape == (1, 2, 32, 32) and torch.isfinite(y).all()
    with pytest.raises(NotImplementedError):
        networks.define_G(SimpleNamespace(net_G="IFNet", n_class=1, precision="tf32"), gpu_ids=[0])
    with pytest.raises(NotImplementedError):
        net.forward_uint8(torch.zeros(1, 32, 32, 3, dtype=torch.uint8).cuda(), torch.zeros(1, 32, 32, 3, dtype=torch.uint8).cuda())
