"""BIT (BASE_Transformer, all three registered depths) and its ResNet-18 baseline on the GPU against the oracle, the
emulator and the golden fixtures generated from the unmodified reference (models/networks.py:223-441): logits within
2e-2 absolute (bf16 path), change maps agreeing on >= 99.9 % of decided pixels."""
import os

import numpy as np
import pytest
import torch

from oracle import emulate, nets
from stcd_b200 import bit, synth

pytestmark = pytest.mark.gpu
BF16_TOL = 2e-2
VARIANTS = {"s4": {}, "dd8": dict(enc_depth=1, dec_depth=8), "dd8_dedim8": dict(enc_depth=1, dec_depth=8, decoder_dim_head=8)}


def _net(variant="dd8"):
    if variant == "resnet18":
        return synth.prepare_(bit.ResNet(3, 2).eval(), "ResNet"), 5
    net = bit.BASE_Transformer(3, 2, with_pos="learned", resnet_stages_num=4, token_len=4, **VARIANTS[variant])
    return synth.prepare_(net.eval(), "BASE_Transformer"), 4


def _out(y):
    return y[0] if isinstance(y, (list, tuple)) else y


def _agreement(y, ref):
    margin = (ref[:, 1] - ref[:, 0]).abs()
    agree = (y[:, 1] > y[:, 0]) == (ref[:, 1] > ref[:, 0])
    return agree.float().mean().item(), agree[margin > BF16_TOL].float().mean().item()


@pytest.mark.parametrize("variant", ["s4", "dd8", "dd8_dedim8", "resnet18"])
def test_forward_matches_oracle_and_emulator(variant):
    net, stages = _net(variant)
    x1, x2 = synth.image_pairs(3, 64, 96)
    with torch.no_grad():
        ref = nets.bit_forward(net.state_dict(), x1, x2, stages=stages)
    emu = emulate.run_program(net.lower(64, 96), x1[:2], x2[:2], chunk=2)[0]
    net = net.cuda()
    net.chunk_pairs = 2                      # 3 pairs -> one full chunk + a ragged one
    y = net(x1.cuda(), x2.cuda())
    assert isinstance(y, list) == (variant != "resnet18"), "BASE_Transformer returns [logits], ResNet the tensor"
    y = _out(y).cpu()
    assert y.shape == ref.shape and y.dtype == torch.float32
    assert (y[:2] - emu).abs().max().item() < 1.5e-2, "kernel vs emulator (same rounding points; bf16 flips cascade)"
    assert (y - ref).abs().max().item() < BF16_TOL, "kernel vs fp32 oracle"
    assert _agreement(y, ref)[1] >= 0.999


def test_token_path_against_emulator():
    """The K12 op alone: plan tensors before / after it against the emulator's (fp32 arithmetic on both sides)."""
    net, _ = _net("dd8")
    x1, x2 = synth.image_pairs(2, 64, 64)
    keep = {}
    emulate.run_program(net.lower(64, 64), x1, x2, chunk=2, keep=keep)
    net = net.cuda()
    net.chunk_pairs = 2
    net(x1.cuda(), x2.cuda())
    torch.cuda.synchronize()
    plan = net.plan_for(x1.cuda())
    for name, lim in (("conv_pred.o", 0.02), ("bit.o", 0.02), ("diff", 0.08), ("classifier.t", 0.05)):
        got, want = plan.read_tensor(name), keep[name]
        err = ((got - want).abs().mean() / (want.abs().mean() + 1e-3)).item()
        assert err < lim, (name, err)


@pytest.mark.parametrize("case", ["bit_dd8", "bit_resnet18"])
def test_forward_matches_golden(case, golden_dir):
    g = np.load(os.path.join(golden_dir, f"{case}.npz"))
    net, _ = _net("dd8" if case == "bit_dd8" else "resnet18")
    net = net.cuda()
    x1, x2 = synth.image_pairs(int(g["batch"]), int(g["h"]), int(g["w"]), seed=int(g["data_seed"]))
    y = _out(net(x1.cuda(), x2.cuda())).cpu()
    ref = torch.from_numpy(g["out0"])
    assert (y - ref).abs().max().item() < BF16_TOL
    assert _agreement(y, ref)[1] >= 0.999


def test_256_shape_and_properties():
    """256x256 (the reference's img_size) at a batch the oracle finishes in seconds, plus determinism, batch-order
    equivariance and host-buffer path == device path."""
    net, stages = _net("dd8")
    x1, x2 = synth.image_pairs(5, 256, 256)
    with torch.no_grad():
        ref = nets.bit_forward(net.state_dict(), x1[:1], x2[:1], stages=stages)
    net = net.cuda()
    net.chunk_pairs = 4
    y = _out(net(x1.cuda(), x2.cuda()))
    assert (y[:1].cpu() - ref).abs().max().item() < BF16_TOL
    frac = (ref[:, 1] > ref[:, 0]).float().mean().item()
    assert 0.02 < frac < 0.98, "degenerate change map"
    assert torch.equal(y, _out(net(x1.cuda(), x2.cuda()))), "forward must be deterministic"
    perm = torch.tensor([3, 1, 4, 0, 2])
    yp = _out(net(x1[perm].cuda(), x2[perm].cuda()))
    assert torch.equal(yp, y[perm.cuda()]), "pairs are independent: permuting the batch permutes the logits"
    plan = net.plan_for(x1.cuda())
    outs = plan.forward_host(x1.pin_memory(), x2.pin_memory())
    assert torch.equal(outs[0], y.cpu()), "host-buffer path must equal the device path bit for bit"


def test_define_G_returns_bit():
    from types import SimpleNamespace
    from stcd_b200 import networks
    for key, cls in (("base_transformer_pos_s4_dd8", bit.BASE_Transformer), ("base_resnet18", bit.ResNet)):
        net = networks.define_G(SimpleNamespace(net_G=key, n_class=2), gpu_ids=[0])
        assert isinstance(net, cls) and next(net.parameters()).is_cuda
