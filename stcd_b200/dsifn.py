"""IFNet / DSIFN (deeply supervised image fusion network) behind the reference's ``net_G(x1, x2)`` contract.

Drop-in for ``models/DSIFN.py::vgg16_base`` (:9-21) and ``DSIFN`` (:63-188; registry key ``IFNet``,
models/networks.py:164-166): same constructors, same parameter names (``t1_base.features.0.weight`` ...,
``o2_conv1.0.weight`` / ``.1.weight`` (PReLU) / ``.2.*`` (BatchNorm), ``sa3.conv1.weight``, ``ca4.fc1.weight`` ...; a
reference ``state_dict`` loads, including the unused ``ca1`` / ``bn_ca*`` / ``o*_conv3|4`` side heads), same return value:
the single-channel logits ``[B, 1, H, W]`` of ``o5_conv4`` (the four deep-supervision sigmoids go to a list the reference
discards, :133,147,159,171, so they are not computed).  Eval mode (Dropout is the identity).  The VGG16 checkpoint is never
downloaded (``pretrained=True`` upstream): load a ``state_dict``.  The registry shares ONE ``vgg16_base`` between the two
dates; two different bases are rejected at lowering time (the encoder runs both dates as Siamese pair tiles).

Lowering:

* VGG16 ``features[:30]``: 13 conv + bias + ReLU launches on both dates at once; the four 2x2 max-pools are fused into the
  producing conv's epilogue (it writes the tap AND the pooled map);
* every ``conv2d_bn`` (conv -> PReLU -> BatchNorm, :54-60) is one conv with the activation and the second affine in its
  epilogue; ``torch.cat`` of the up-sampled map with the two dates' features is virtual for the first branch and
  materialised (already scaled) by the channel-attention op for the others;
* ``ca(x) * x`` and ``bn(sa(x) * x)`` are the K13 bandwidth kernels; ``ConvTranspose2d(k=2, s=2)`` is 4 single-tap phases.
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch
import torch.nn as nn

from . import lowering as L
from .module import PlannedModule

_VGG_CFG = (64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "M", 512, 512, 512)     # vgg16().features[:30]
_TAPS = (3, 8, 15, 22, 29)


class vgg16_base(nn.Module):
    """models/DSIFN.py:9-21 (parameters only; torchvision's layer indices)."""

    def __init__(self):
        super().__init__()
        layers, cin = [], 3
        for v in _VGG_CFG:
            if v == "M":
                layers.append(nn.MaxPool2d(kernel_size=2, stride=2))
            else:
                layers += [nn.Conv2d(cin, v, 3, padding=1), nn.ReLU(inplace=True)]
                cin = v
        assert len(layers) == 30
        self.features = nn.ModuleList(layers).eval()


class _ChannelAttention(nn.Module):
    """models/DSIFN.py:24-31."""

    def __init__(self, in_channels: int, ratio: int = 8):
        super().__init__()
        self.fc1 = nn.Conv2d(in_channels, in_channels // ratio, 1, bias=False)
        self.fc2 = nn.Conv2d(in_channels // ratio, in_channels, 1, bias=False)


class _SpatialAttention(nn.Module):
    """models/DSIFN.py:39-43."""

    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(2, 1, 7, padding=3, bias=False)


def _conv2d_bn(cin: int, cout: int) -> nn.Sequential:
    """models/DSIFN.py:54-60."""
    return nn.Sequential(nn.Conv2d(cin, cout, 3, padding=1), nn.PReLU(), nn.BatchNorm2d(cout), nn.Dropout(p=0.6))


class DSIFN(PlannedModule):
    """models/DSIFN.py:63-188."""
    default_chunk_pairs = 8

    def __init__(self, model_A: vgg16_base, model_B: vgg16_base):
        super().__init__()
        self.t1_base = model_A
        self.t2_base = model_B
        for i in range(1, 6):
            setattr(self, f"sa{i}", _SpatialAttention())
        # branch 1 (:77-84)
        self.ca1 = _ChannelAttention(1024)
        self.bn_ca1 = nn.BatchNorm2d(1024)
        self.o1_conv1 = _conv2d_bn(1024, 512)
        self.o1_conv2 = _conv2d_bn(512, 512)
        self.bn_sa1 = nn.BatchNorm2d(512)
        self.o1_conv3 = nn.Conv2d(512, 1, 1)
        self.trans_conv1 = nn.ConvTranspose2d(512, 512, kernel_size=2, stride=2)
        # branch 2 (:86-94)
        self.ca2 = _ChannelAttention(1536)
        self.bn_ca2 = nn.BatchNorm2d(1536)
        self.o2_conv1 = _conv2d_bn(1536, 512)
        self.o2_conv2 = _conv2d_bn(512, 256)
        self.o2_conv3 = _conv2d_bn(256, 256)
        self.bn_sa2 = nn.BatchNorm2d(256)
        self.o2_conv4 = nn.Conv2d(256, 1, 1)
        self.trans_conv2 = nn.ConvTranspose2d(256, 256, kernel_size=2, stride=2)
        # branches 3-5 (:96-117)
        for b, (cin, c1, c2, c3) in ((3, (768, 256, 128, 128)), (4, (384, 128, 64, 64)), (5, (192, 64, 32, 16))):
            setattr(self, f"ca{b}", _ChannelAttention(cin))
            setattr(self, f"o{b}_conv1", _conv2d_bn(cin, c1))
            setattr(self, f"o{b}_conv2", _conv2d_bn(c1, c2))
            setattr(self, f"o{b}_conv3", _conv2d_bn(c2, c3))
            setattr(self, f"bn_sa{b}", nn.BatchNorm2d(c3))
            setattr(self, f"o{b}_conv4", nn.Conv2d(c3, 1, 1))
            if b < 5:
                setattr(self, f"trans_conv{b}", nn.ConvTranspose2d(c3, c3, kernel_size=2, stride=2))

    def lower(self, h: int, w: int) -> L.Program:
        return lower_dsifn(self.state_dict(), h, w)

    @torch.no_grad()
    def forward(self, t1_input: torch.Tensor, t2_input: torch.Tensor) -> torch.Tensor:
        return self.plan_for(t1_input).forward(t1_input, t2_input)[0]


# ------------------------------------------------------------------------------------------
def lower_dsifn(sd: Dict[str, torch.Tensor], h: int, w: int) -> L.Program:
    """state_dict of the reference DSIFN -> fused-op Program (eval mode)."""
    if h % 16 or w % 16:
        raise ValueError(f"DSIFN lowering needs H and W divisible by 16 (got {h}x{w}): the reference's skip concats do not line up otherwise")
    sd = {k: v.detach().to("cpu", torch.float32) for k, v in sd.items()}
    for k, v in sd.items():
        if k.startswith("t1_base.") and not torch.equal(v, sd["t2_base." + k[len("t1_base."):]]):
            raise NotImplementedError("stcd_b200.DSIFN runs the two dates through ONE shared VGG16 (DSIFN(base, base), "
                                      f"models/networks.py:164-166); t1_base and t2_base differ at {k}")
    p = L.Program(model="DSIFN", in_channels=3, h=h, w=w)
    ones = lambda c: np.ones(c, np.float32)  # noqa: E731

    def f32(t: torch.Tensor) -> np.ndarray:
        return np.ascontiguousarray(t.numpy().astype(np.float32))

    # ---------------- shared VGG16 features[:30] on both dates (:120-121)
    p.tensor("in", 2, h, w, 8)
    p.ops.append(L.InputPackSpec("pack", "in", 3))
    x, cin, hh, ww = "in", 3, h, w
    feats = []
    conv_ids = [i for i in range(30) if f"t1_base.features.{i}.weight" in sd]
    for i in conv_ids:
        wt = sd[f"t1_base.features.{i}.weight"]
        cout = wt.shape[0]
        tap = (i + 1) in _TAPS
        pooled = tap and (i + 1) != 29                       # features[i + 2] is the 2x2 max-pool
        o = p.tensor(f"vgg.{i}", 2, hh, ww, cout)
        op_ = p.tensor(f"vgg.{i}.pool", 2, hh // 2, ww // 2, cout) if pooled else None
        L.add_conv(p, f"t1_base.features.{i}", [L.Segment(x, cin)], L.conv_taps(wt, pad=1), cout, hh, ww, 1, ones(cout),
                   f32(sd[f"t1_base.features.{i}.bias"]), pair=True, relu=True, out0=o, out_pool=op_,
                   macs_per_pair=2 * hh * ww * 9 * cin * cout)
        if tap:
            feats.append((o, cout))
        x, cin = (op_, cout) if pooled else (o, cout)
        if pooled:
            hh, ww = hh // 2, ww // 2

    def conv_bn(name: str, segs, c_in: int, c_out: int, hh_: int, ww_: int) -> str:
        s2, b2 = L.fold_bn(None, L.bn_params(sd, f"{name}.2"), c_out)
        o_ = p.tensor(f"{name}.o", 1, hh_, ww_, (c_out + 7) // 8 * 8)
        L.add_conv(p, name, segs, L.conv_taps(sd[f"{name}.0.weight"], pad=1), c_out, hh_, ww_, 1, ones(c_out), f32(sd[f"{name}.0.bias"]),
                   act="prelu", act_alpha=float(sd[f"{name}.1.weight"][0]), act_pre=True, scale2=s2, shift2=b2, out0=o_,
                   macs_per_pair=hh_ * ww_ * 9 * c_in * c_out)
        return o_

    def spatial_gate(b: int, src: str, c: int, hh_: int, ww_: int) -> str:
        sc, sh = L.fold_bn(None, L.bn_params(sd, f"bn_sa{b}"), c)
        o_ = p.tensor(f"sa{b}.o", 1, hh_, ww_, c)
        p.ops.append(L.SpatialGateSpec(f"sa{b}", src, o_, c, f32(sd[f"sa{b}.conv1.weight"][0]), sc, sh))
        return o_

    # ---------------- branch 1 (:126-132): cat(t1_l29, t2_l29) is virtual
    f29, c29 = feats[4]
    x = conv_bn("o1_conv1", [L.Segment(f29, c29, stream=0), L.Segment(f29, c29, stream=1)], 2 * c29, 512, hh, ww)
    x = conv_bn("o1_conv2", [L.Segment(x, 512)], 512, 512, hh, ww)
    x = spatial_gate(1, x, 512, hh, ww)
    cx = 512
    # ---------------- branches 2-5 (:135-183)
    for b in (2, 3, 4, 5):
        wt = sd[f"trans_conv{b - 1}.weight"]
        up = p.tensor(f"trans_conv{b - 1}.o", 1, 2 * hh, 2 * ww, cx)
        L.add_conv(p, f"trans_conv{b - 1}", [L.Segment(x, cx)], L.convT_phase_taps(wt, 2, 0), cx, hh, ww, 1, ones(cx),
                   f32(sd[f"trans_conv{b - 1}.bias"]), osy=2, osx=2, out0=up, macs_per_pair=hh * ww * 4 * cx * cx)
        hh, ww = 2 * hh, 2 * ww
        ft, fc = feats[5 - b]
        ctot = cx + 2 * fc
        cat = p.tensor(f"ca{b}.o", 1, hh, ww, ctot)
        hid = sd[f"ca{b}.fc1.weight"].shape[0]
        p.ops.append(L.ChannelAttentionSpec(f"ca{b}", [(up, 0, cx), (ft, 0, fc), (ft, 1, fc)], cat,
                                            f32(sd[f"ca{b}.fc1.weight"].reshape(hid, ctot)), f32(sd[f"ca{b}.fc2.weight"].reshape(ctot, hid))))
        x, cx = cat, ctot
        for i in (1, 2, 3):
            cout = sd[f"o{b}_conv{i}.0.weight"].shape[0]
            x = conv_bn(f"o{b}_conv{i}", [L.Segment(x, cx)], cx, cout, hh, ww)
            cx = cout
        x = spatial_gate(b, x, cx, hh, ww)
    L.add_conv(p, "o5_conv4", [L.Segment(x, cx)], L.conv_taps(sd["o5_conv4.weight"], pad=0), 1, hh, ww, 1, ones(1), f32(sd["o5_conv4.bias"]),
               out_ext=0, macs_per_pair=hh * ww * cx)
    p.ext.append(L.ExtOutput("out", 1, hh, ww))
    return p
