"""SNUNet-CD (ECAM) on the GPU against the oracle, the emulator and the golden fixture generated from
the unmodified reference (models/SNUNet.py): logits within 2e-2 absolute (bf16 path), change maps
agreeing on >= 99.9 % of decided pixels."""
import os

import numpy as np
import pytest
import torch

import parity
from oracle import emulate, nets
from stcd_b200 import snunet, synth

pytestmark = pytest.mark.gpu
BF16_TOL = 2e-2


def _net():
    return synth.prepare_(snunet.SNUNet_ECAM(3, 2).eval(), "SNUNet_ECAM")


def _agreement(y, ref):
    margin = (ref[:, 1] - ref[:, 0]).abs()
    agree = (y[:, 1] > y[:, 0]) == (ref[:, 1] > ref[:, 0])
    return agree.float().mean().item(), agree[margin > BF16_TOL].float().mean().item()


def test_forward_matches_oracle_and_emulator():
    net = _net()
    x1, x2 = synth.image_pairs(3, 64, 96)
    with torch.no_grad():
        ref = nets.snunet_forward(net.state_dict(), x1, x2)
    emu = emulate.run_program(net.lower(64, 96), x1, x2, chunk=2)[0]
    net = net.cuda()
    net.chunk_pairs = 2                      # 3 pairs -> one full chunk + a ragged one
    y = net(x1.cuda(), x2.cuda())
    assert isinstance(y, torch.Tensor) and y.shape == ref.shape and y.dtype == torch.float32
    y = y.cpu()
    assert (y - emu).abs().max().item() < 1.5e-2, "kernel vs emulator (same rounding points; bf16 flips cascade)"
    assert (y - ref).abs().max().item() < BF16_TOL, "kernel vs fp32 oracle"
    parity.check("snunet:3x64x96", y, ref, "argmax")       # absolute + relative criteria, logit std >= 0.25 (tests/parity.py)


def test_forward_matches_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "snunet.npz"))
    net = synth.prepare_(snunet.SNUNet_ECAM(3, 2).eval(), "SNUNet_ECAM", seed=int(g["weight_seed"])).cuda()
    x1, x2 = synth.image_pairs(int(g["batch"]), int(g["h"]), int(g["w"]), seed=int(g["data_seed"]))
    y = net(x1.cuda(), x2.cuda()).cpu()
    ref = torch.from_numpy(g["out0"])
    assert (y - ref).abs().max().item() < BF16_TOL
    assert _agreement(y, ref)[1] >= 0.999
    parity.check("golden:snunet", y, ref, "argmax", all_px=0.99)


def test_config_c2_full_batch_chunk64_vs_oracle():
    """Config C2 exactly as bench.py runs it: 64 pairs in ONE chunk of 64 (the widest tile walk: MT sub-tiles span different
    images of the chunk), compared with the fp32 oracle on four pairs drawn from different positions of the chunk."""
    net = _net()
    x1, x2 = synth.image_pairs(64, 256, 256)
    idx = [0, 21, 42, 63]
    with torch.no_grad():
        ref = nets.snunet_forward(net.state_dict(), x1[idx], x2[idx])
    net = net.cuda()
    net.chunk_pairs = 64
    y = net(x1.cuda(), x2.cuda())
    assert y.shape == (64, 2, 256, 256)
    parity.check("snunet:C2 64x256x256 chunk 64 (pairs 0,21,42,63 vs oracle)", y[idx], ref, "argmax")
    # the chunk walk must not matter: the same pairs through chunks of 16 give the same logits bit for bit
    net.chunk_pairs = 16
    y16 = net(x1.cuda(), x2.cuda())
    assert torch.equal(y16, y), "logits must not depend on the chunk size"


def test_config_c2_shape_and_properties():
    """Config C2 shape (256x256) at a batch the oracle finishes in seconds, plus the size-independent
    properties: determinism, batch-order equivariance, host-buffer path == device path."""
    net = _net()
    x1, x2 = synth.image_pairs(6, 256, 256)
    with torch.no_grad():
        ref = nets.snunet_forward(net.state_dict(), x1[:1], x2[:1])
    net = net.cuda()
    net.chunk_pairs = 4
    y = net(x1.cuda(), x2.cuda())
    assert (y[:1].cpu() - ref).abs().max().item() < BF16_TOL
    parity.check("snunet:6x256x256 chunk 4 (pair 0 vs oracle)", y[:1], ref, "argmax")
    assert torch.equal(y, net(x1.cuda(), x2.cuda())), "forward must be deterministic"
    perm = torch.tensor([3, 1, 5, 0, 2, 4])
    yp = net(x1[perm].cuda(), x2[perm].cuda())
    assert torch.equal(yp, y[perm.cuda()]), "pairs are independent: permuting the batch permutes the logits"
    plan = net.plan_for(x1.cuda())
    outs = plan.forward_host(x1.pin_memory(), x2.pin_memory())
    assert torch.equal(outs[0], y.cpu()), "host-buffer path must equal the device path bit for bit"
    assert plan.launches(4) == 43            # pack + 40 convs + 2 ECAM kernels


def test_define_G_returns_snunet():
    from types import SimpleNamespace
    from stcd_b200 import networks
    net = networks.define_G(SimpleNamespace(net_G="SNUNet", n_class=2), gpu_ids=[0])
    assert isinstance(net, snunet.SNUNet_ECAM) and next(net.parameters()).is_cuda
    x1, x2 = synth.image_pairs(1, 32, 32)
    y = net.eval()(x1.cuda(), x2.cuda())
    assert y.shape == (1, 2, 32, 32) and torch.isfinite(y).all()
