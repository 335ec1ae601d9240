"""uint8 input pipeline: the decoded uint8 HWC images go straight to the GPU and the pack
kernel applies the reference loader's ToTensor + Normalize (data/dataset.py:196-203).  The bar is bit-exact:
``net.forward_uint8(a, b) == net(normalize(a), normalize(b))`` for every net family, plus the raw {0, 255} mask label
binarised on the fly (data/dataset.py:206-210)."""
import numpy as np
import pytest
import torch

from oracle import emulate, nets
from stcd_b200 import segcd, siamunet, snunet, synth
from stcd_b200.metric import SegmentationMetric
from stcd_b200.module import PlannedModule

pytestmark = pytest.mark.gpu
MEAN, STD = PlannedModule.IMAGENET_MEAN, PlannedModule.IMAGENET_STD


def _images(b, h, w, seed=21):
    g = torch.Generator().manual_seed(seed)
    a = torch.randint(0, 256, (b, h, w, 3), generator=g, dtype=torch.uint8)
    noise = torch.randint(-40, 41, (b, h, w, 3), generator=g)
    bb = (a.to(torch.int64) + noise).clamp(0, 255).to(torch.uint8)
    return a, bb


@pytest.mark.parametrize("family", ["SiamUnet_diff", "SNUNet_ECAM", "SegCD"])
def test_uint8_forward_is_bit_identical_to_host_normalised_forward(family):
    net = {"SiamUnet_diff": lambda: siamunet.SiamUnet_diff(3, 2), "SNUNet_ECAM": lambda: snunet.SNUNet_ECAM(3, 2),
           "SegCD": lambda: segcd.SegCD("resnet34")}[family]()
    net = synth.prepare_(net.eval(), family).cuda()
    net.chunk_pairs = 2
    a, b = _images(3, 64, 96)
    xa, xb = emulate.normalize_u8(a, MEAN, STD), emulate.normalize_u8(b, MEAN, STD)
    want = net(xa.cuda(), xb.cuda())
    got = net.forward_uint8(a.cuda(), b.cuda())
    want = want if isinstance(want, tuple) else (want,)
    got = got if isinstance(got, tuple) else (got,)
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert torch.equal(g, w)
    # and the host-buffer path
    plan = net.plan_for(a.cuda(), u8_norm=(MEAN, STD))
    outs = plan.forward_host(a.pin_memory(), b.pin_memory())
    for o, w in zip(outs, want):
        assert torch.equal(o, w.cpu())


def test_uint8_forward_matches_oracle():
    net = synth.prepare_(siamunet.SiamUnet_diff(3, 2).eval(), "SiamUnet_diff")
    a, b = _images(2, 48, 64, seed=5)
    with torch.no_grad():
        ref = nets.siamunet_forward(net.state_dict(), emulate.normalize_u8(a, MEAN, STD), emulate.normalize_u8(b, MEAN, STD), "diff")
    y = net.cuda().forward_uint8(a.cuda(), b.cuda()).cpu()
    assert (y - ref).abs().max().item() < 2e-2


def test_uint8_errors_and_raw_mask_label():
    net = synth.prepare_(siamunet.SiamUnet_diff(3, 2).eval(), "SiamUnet_diff").cuda()
    a, b = _images(2, 32, 32)
    with pytest.raises(TypeError):
        net.forward_uint8(a.cuda().float(), b.cuda().float())
    with pytest.raises(ValueError):
        net.forward_uint8(a.cuda().permute(0, 3, 1, 2).contiguous(), b.cuda().permute(0, 3, 1, 2).contiguous())
    y = net.forward_uint8(a.cuda(), b.cuda())
    g = torch.Generator().manual_seed(9)
    mask = ((torch.rand(2, 32, 32, generator=g) < 0.3).to(torch.uint8) * 255)          # PNG mask as stored on disk
    m = SegmentationMetric(2, "cuda:0")
    m.addLogits(y, mask.cuda(), kind="argmax", raw_mask_label=True)
    label = (mask >= 1).long()
    pred = y.cpu().argmax(1)
    want = np.bincount((2 * label + pred).reshape(-1).numpy(), minlength=4).reshape(2, 2)
    assert np.array_equal(m.confusion_counts().cpu().numpy(), want)
