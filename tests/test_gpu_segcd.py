"""smp.SegCD(resnet34)(A, B) on the GPU against the oracle, the emulator and the golden fixture generated
from the unmodified reference (north-star tolerances: outputs within 2e-2 absolute on the bf16 path,
binary change maps agreeing on >= 99.9 % of decided pixels)."""
import os

import numpy as np
import pytest
import torch

import parity
from oracle import emulate, nets
from stcd_b200 import segcd, synth
from stcd_b200.metric import SegmentationMetric

pytestmark = pytest.mark.gpu
BF16_TOL = 2e-2


def _net(name="resnet34"):
    return synth.prepare_(segcd.SegCD(name).eval(), "SegCD")


def _check(ys, refs, tag="segcd"):
    """The masks m1 / m2 are held to the shared relative criteria (tests/parity.py); the logit spread is NOT raised to 0.25 for
    this family: ~50 fused bf16 layers put the rms error at ~2 % of the spread (oracle/emulate.py reproduces it on the CPU),
    so 2e-2 absolute only holds up to a spread of ~0.15, and `change` = min(head(|d1 - d2|), |m1 - m2|) is a DIFFERENCE of two
    nearly equal Siamese outputs whose error is that of the masks against a spread half as wide.  Users who need more take
    the split-precision path (precision="tf32")."""
    assert len(ys) == 3
    for k, (y, ref) in enumerate(zip(ys, refs)):
        assert y.shape == ref.shape and y.dtype == torch.float32
        assert (y.cpu() - ref).abs().max().item() < BF16_TOL
        r = parity.report(f"{tag}:{('m1', 'm2', 'change')[k]}", y, ref, "sigmoid")
        if k < 2:
            assert r["rms_over_std"] <= 0.03 and r["max_over_std"] <= 0.2, r      # the tail over up to 4 * 10**6 logits
            assert r["agree_decided"] >= 0.999 and r["agree_all"] >= 0.99, r
    change, ref = ys[2].cpu(), refs[2]
    agree = (change > 0) == (ref > 0)
    assert agree[ref.abs() > BF16_TOL].float().mean().item() >= 0.999
    assert agree.float().mean().item() >= 0.97


@pytest.mark.parametrize("name", ["resnet34", "resnet18", "resnet50"])
def test_forward_matches_oracle_and_emulator(name):
    net = _net(name)
    x1, x2 = synth.image_pairs(3, 64, 96)
    layers = segcd._LAYERS[name]
    with torch.no_grad():
        ref = nets.segcd_forward(net.state_dict(), x1, x2, layers)
    emu = emulate.run_program(net.lower(64, 96), x1, x2, chunk=2)
    net = net.cuda()
    net.chunk_pairs = 2                      # 3 pairs -> one full chunk + a ragged one
    ys = net(x1.cuda(), x2.cuda())
    assert isinstance(ys, tuple)
    _check(ys, ref)
    for y, e in zip(ys, emu):
        assert (y.cpu() - e).abs().max().item() < 1.5e-2, "kernel vs emulator (same rounding points)"


@pytest.mark.parametrize("case,name", [("segcd_r34", "resnet34"), ("segcd_r50", "resnet50")])
def test_forward_matches_golden(golden_dir, case, name):
    g = np.load(os.path.join(golden_dir, f"{case}.npz"))
    assert float(g["gain"]) == synth.GAINS["SegCD"] and int(g["n_out"]) == 3
    net = _net(name).cuda()
    x1, x2 = synth.image_pairs(int(g["batch"]), int(g["h"]), int(g["w"]), seed=int(g["data_seed"]))
    ys = net(x1.cuda(), x2.cuda())
    _check(ys, [torch.from_numpy(g[f"out{i}"]) for i in range(3)])


def test_ffctlcd_matches_oracle_and_emulator():
    """smp.FFCTLCD (decoders/unet/model.py:335-423): third decoder pass over |f1 - f2| of every encoder feature."""
    net = synth.prepare_(segcd.FFCTLCD("resnet34").eval(), "SegCD")
    x1, x2 = synth.image_pairs(3, 64, 96)
    with torch.no_grad():
        ref = nets.ffctlcd_forward(net.state_dict(), x1, x2)
    emu = emulate.run_program(net.lower(64, 96), x1, x2, chunk=2)
    net = net.cuda()
    net.chunk_pairs = 2
    ys = net(x1.cuda(), x2.cuda())
    _check(ys, ref)
    for y, e in zip(ys, emu):
        assert (y.cpu() - e).abs().max().item() < 1.5e-2
    import stcd_b200.smp as smp
    assert isinstance(smp.create_model("FFCTLCD", "resnet18"), segcd.FFCTLCD)


def test_layerwise_against_emulator():
    """Every intermediate tensor of the plan against the emulator: localises a wrong layer."""
    net = _net()
    x1, x2 = synth.image_pairs(2, 64, 64)
    keep = {}
    emulate.run_program(net.lower(64, 64), x1, x2, chunk=2, keep=keep)
    net = net.cuda()
    net.chunk_pairs = 2
    net(x1.cuda(), x2.cuda())
    torch.cuda.synchronize()
    plan = net.plan_for(x1.cuda())
    worst = {}
    for name, t in plan.prog.tensors.items():
        got = plan.read_tensor(name)
        # same rounding points, different fp32 accumulation order: isolated bf16-ulp flips that cascade
        # through ~40 layers.  A wrong layer is O(1) off everywhere; drift stays below 1 % in the mean.
        worst[name] = ((got - keep[name]).abs().mean() / (keep[name].abs().mean() + 1e-6)).item()
    bad = {k: v for k, v in worst.items() if v > 1e-2}
    assert not bad, bad


def test_config_c3_full_batch_chunk16_vs_oracle():
    """Config C3 exactly as bench.py runs it: 16 tiles of 1024x1024 in ONE chunk of 16, compared with the fp32 oracle on four
    pairs drawn from different positions of the chunk."""
    net = _net()
    x1, x2 = synth.image_pairs(16, 1024, 1024)
    idx = [0, 5, 10, 15]
    with torch.no_grad():
        ref = nets.segcd_forward(net.state_dict(), x1[idx], x2[idx])
    net = net.cuda()
    net.chunk_pairs = 16
    ys = net(x1.cuda(), x2.cuda())
    assert ys[2].shape == (16, 1, 1024, 1024)
    _check([y[idx] for y in ys], ref, tag="segcd_r34:C3 16x1024x1024 chunk 16 (pairs 0,5,10,15 vs oracle)")
    net.chunk_pairs = 4
    ys4 = net(x1.cuda(), x2.cuda())
    assert all(torch.equal(u, v) for u, v in zip(ys4, ys)), "outputs must not depend on the chunk size"


def test_full_size_tile_and_properties():
    """Config C3's tile size (1024x1024), 2 pairs: the first pair against the CPU oracle, then the
    size-independent properties (determinism, batch-order equivariance, identical images -> no change)
    and the evaluator fed by sigmoid(change) > 0.5 (train_stcd.py:477-492)."""
    net = _net()
    x1, x2 = synth.image_pairs(2, 1024, 1024)
    with torch.no_grad():
        ref = nets.segcd_forward(net.state_dict(), x1[:1], x2[:1])
    net = net.cuda()
    net.chunk_pairs = 1
    a, b = x1.cuda(), x2.cuda()
    ys = net(a, b)
    _check([y[:1] for y in ys], ref)
    ys2 = net(a, b)
    assert all(torch.equal(u, v) for u, v in zip(ys, ys2)), "forward must be deterministic"
    perm = torch.tensor([1, 0])
    yp = net(a[perm], b[perm])
    assert all(torch.equal(u, v[perm.cuda()]) for u, v in zip(yp, ys))
    same = net(a, a)
    assert torch.equal(same[0], same[1]), "shared weights: identical images give identical masks"
    # |m1 - m2| == 0 and head(0) == bias: change = min(bias, 0)
    bias = float(net.segmentation_head[0].bias[0])
    assert torch.allclose(same[2], torch.full_like(same[2], min(bias, 0.0)), atol=1e-6)
    label = synth.labels(2, 1024, 1024).cuda()
    m = SegmentationMetric(2, "cuda:0")
    m.addLogits(ys[2], label, kind="sigmoid", thr=0.5)
    cm = m.confusion_counts().cpu().numpy()
    pred = (torch.sigmoid(ys[2].cpu()) > 0.5).long()[:, 0]
    want = np.bincount((2 * label.cpu() + pred).reshape(-1).numpy(), minlength=4).reshape(2, 2)
    assert np.array_equal(cm, want) and cm.sum() == 2 * 1024 * 1024
    plan = net.plan_for(a)
    outs = plan.forward_host(x1.pin_memory(), x2.pin_memory())
    assert all(torch.equal(o, y.cpu()) for o, y in zip(outs, ys)), "host-buffer path must equal the device path"


def test_reference_calling_conventions():
    import stcd_b200.smp as smp
    net = smp.create_model("SegCD", "resnet34", None, classes=1).eval()
    ref = _net()
    net.load_state_dict(ref.state_dict())          # reference parameter names
    x1, x2 = synth.image_pairs(1, 32, 64)
    m1, m2, change = net.cuda()(x1.cuda(), x2.cuda())      # train_stcd.py:476 unpacks three
    r = ref.cuda()(x1.cuda(), x2.cuda())
    assert torch.equal(m1, r[0]) and torch.equal(change, r[2])
    with pytest.raises(KeyError):
        smp.create_model("nope")
    with pytest.raises(ValueError):
        net(torch.zeros(1, 3, 40, 40).cuda(), torch.zeros(1, 3, 40, 40).cuda())
