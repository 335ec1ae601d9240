"""SNUNet-CD (ECAM) behind the reference's ``net_G(xA, xB)`` contract.

Drop-in for ``models/SNUNet.py::SNUNet_ECAM`` (registry key ``"SNUNet"``, models/networks.py:168-169):
same ctor arguments, same parameter names, same return type (one ``[B, out_ch, H, W]`` tensor).
Eval-mode lowering to libstcd_b200 (one launch per line):

* ``conv_block_nested`` (SNUNet.py:8-26) = two implicit-GEMM convs: conv1 writes the PRE-BatchNorm
  output (the block's residual, :18-19) and the BN1+ReLU'd one from the same accumulator; conv2
  folds BN2, adds the residual, applies ReLU and — for encoder nodes — also writes the 2x2
  max-pooled tensor the next level reads (:120-129).
* encoder nodes run both temporal images per launch (Siamese pair tiles); ``conv4_0`` runs on
  image B only, as upstream (:123 is commented out, :129).
* ``up`` = ConvTranspose2d(C, C, 2, stride=2) (:38) runs as 4 single-tap output phases.
* the dense skip ``torch.cat`` (:131-142, up to 6 tensors / 224 channels) is never materialised:
  nested-block convs read their sources as K-segments.
* the ECAM tail (:144-149) is one fused op (lowering.EcamHeadSpec): a per-(image, channel) avg/max
  reduction, then a per-image 1x1 head.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn as nn

from . import lowering as L
from .module import PlannedModule

FILTERS = [32, 64, 128, 256, 512]


class conv_block_nested(nn.Module):
    """Parameter holder with the reference's names (SNUNet.py:8-15)."""

    def __init__(self, in_ch: int, mid_ch: int, out_ch: int):
        super().__init__()
        self.conv1 = nn.Conv2d(in_ch, mid_ch, kernel_size=3, padding=1, bias=True)
        self.bn1 = nn.BatchNorm2d(mid_ch)
        self.conv2 = nn.Conv2d(mid_ch, out_ch, kernel_size=3, padding=1, bias=True)
        self.bn2 = nn.BatchNorm2d(out_ch)


class up(nn.Module):  # noqa: N801 - the reference's class name (SNUNet.py:29-43)
    def __init__(self, in_ch: int, bilinear: bool = False):
        super().__init__()
        if bilinear:
            raise NotImplementedError("SNUNet_ECAM constructs up(in_ch) with bilinear=False (SNUNet.py:74-100)")
        self.up = nn.ConvTranspose2d(in_ch, in_ch, 2, stride=2)


class ChannelAttention(nn.Module):
    """SNUNet.py:46-59 (parameters only; the arithmetic runs in the fused ECAM op)."""

    def __init__(self, in_channels: int, ratio: int = 16):
        super().__init__()
        self.fc1 = nn.Conv2d(in_channels, in_channels // ratio, 1, bias=False)
        self.fc2 = nn.Conv2d(in_channels // ratio, in_channels, 1, bias=False)


class SNUNet_ECAM(PlannedModule):
    """models/SNUNet.py:60-152."""
    supports_precision_path = True

    def __init__(self, in_ch: int = 3, out_ch: int = 1):
        super().__init__()
        self.in_ch, self.out_ch = in_ch, out_ch
        f = FILTERS
        self.conv0_0 = conv_block_nested(in_ch, f[0], f[0])
        self.conv1_0 = conv_block_nested(f[0], f[1], f[1])
        self.Up1_0 = up(f[1])
        self.conv2_0 = conv_block_nested(f[1], f[2], f[2])
        self.Up2_0 = up(f[2])
        self.conv3_0 = conv_block_nested(f[2], f[3], f[3])
        self.Up3_0 = up(f[3])
        self.conv4_0 = conv_block_nested(f[3], f[4], f[4])
        self.Up4_0 = up(f[4])
        self.conv0_1 = conv_block_nested(f[0] * 2 + f[1], f[0], f[0])
        self.conv1_1 = conv_block_nested(f[1] * 2 + f[2], f[1], f[1])
        self.Up1_1 = up(f[1])
        self.conv2_1 = conv_block_nested(f[2] * 2 + f[3], f[2], f[2])
        self.Up2_1 = up(f[2])
        self.conv3_1 = conv_block_nested(f[3] * 2 + f[4], f[3], f[3])
        self.Up3_1 = up(f[3])
        self.conv0_2 = conv_block_nested(f[0] * 3 + f[1], f[0], f[0])
        self.conv1_2 = conv_block_nested(f[1] * 3 + f[2], f[1], f[1])
        self.Up1_2 = up(f[1])
        self.conv2_2 = conv_block_nested(f[2] * 3 + f[3], f[2], f[2])
        self.Up2_2 = up(f[2])
        self.conv0_3 = conv_block_nested(f[0] * 4 + f[1], f[0], f[0])
        self.conv1_3 = conv_block_nested(f[1] * 4 + f[2], f[1], f[1])
        self.Up1_3 = up(f[1])
        self.conv0_4 = conv_block_nested(f[0] * 5 + f[1], f[0], f[0])
        self.ca = ChannelAttention(f[0] * 4, ratio=16)
        self.ca1 = ChannelAttention(f[0], ratio=16 // 4)
        self.conv_final = nn.Conv2d(f[0] * 4, out_ch, kernel_size=1)
        for m in self.modules():                     # SNUNet.py:108-113
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def lower(self, h: int, w: int) -> L.Program:
        return lower_snunet(self.state_dict(), self.in_ch, self.out_ch, h, w, precision=self.plan_precision)

    @torch.no_grad()
    def forward(self, xA: torch.Tensor, xB: torch.Tensor) -> torch.Tensor:
        return self.plan_for(xA).forward(xA, xB)[0]


# ------------------------------------------------------------------------------------------
def lower_snunet(sd: Dict[str, torch.Tensor], in_ch: int, out_ch: int, h: int, w: int, precision: str = "bf16") -> L.Program:
    """state_dict of the reference SNUNet_ECAM -> fused-op Program (eval mode)."""
    if h % 16 or w % 16:
        raise ValueError(f"SNUNet lowering needs H and W divisible by 16 (got {h}x{w}): four 2x2 poolings")
    if in_ch > 8:
        raise ValueError("in_ch > 8 not supported by the input packer")
    if out_ch > 4:
        raise ValueError("out_ch > 4 not supported by the fused ECAM head")
    sd = {k: v.detach().to("cpu", torch.float32) for k, v in sd.items()}
    f = FILTERS
    p = L.Program(model="SNUNet_ECAM", in_channels=in_ch, h=h, w=w, precision=precision)
    p.tensor("in", 2, h, w, 8)
    p.ops.append(L.InputPackSpec("pack", "in", in_ch, split=p.split))

    def nested(name: str, segs: Sequence[L.Segment], cout: int, hh: int, ww: int, *, pair: bool,
               pool: Optional[str] = None) -> str:
        """conv_block_nested.forward (SNUNet.py:17-26); returns the output tensor name."""
        mult = 2 if pair else 1
        cin = sum(s.c_real for s in segs)
        w1, b1 = sd[f"{name}.conv1.weight"], sd[f"{name}.conv1.bias"]
        w2, b2 = sd[f"{name}.conv2.weight"], sd[f"{name}.conv2.bias"]
        assert w1.shape[1] == cin, (name, w1.shape, cin)
        ident = p.tensor(f"{name}.id", mult, hh, ww, cout)
        y = p.tensor(f"{name}.y", mult, hh, ww, cout)
        out = p.tensor(f"{name}.out", mult, hh, ww, cout)
        s1, t1 = L.fold_bn(None, L.bn_params(sd, f"{name}.bn1"), cout)        # BN1 applied to (conv1 + bias)
        L.add_conv(p, f"{name}.conv1", segs, L.conv_taps(w1, pad=1), cout, hh, ww, 1,
                   np.ones(cout, np.float32), b1.numpy().astype(np.float32), pair=pair, relu=True,
                   scale2=s1, shift2=t1, out_raw=ident, out0=y, macs_per_pair=mult * hh * ww * 9 * cin * cout)
        s2, t2 = L.fold_bn(b2, L.bn_params(sd, f"{name}.bn2"), cout)
        L.add_conv(p, f"{name}.conv2", [L.Segment(y, cout)], L.conv_taps(w2, pad=1), cout, hh, ww, 1, s2, t2,
                   pair=pair, relu=True, res=ident, out0=out, out_pool=pool,
                   macs_per_pair=mult * hh * ww * 9 * cout * cout)
        return out

    def upsample(name: str, src: str, c: int, hh: int, ww: int, stream: int) -> str:
        """up.forward (SNUNet.py:40-43): ConvTranspose2d(c, c, 2, stride=2) of a [hh, ww] tensor."""
        out = p.tensor(f"{name}.out", 1, 2 * hh, 2 * ww, c)
        L.add_conv(p, name, [L.Segment(src, c, stream=stream)], L.convT_phase_taps(sd[f"{name}.up.weight"], 2, 0), c,
                   hh, ww, 1, np.ones(c, np.float32), sd[f"{name}.up.bias"].numpy().astype(np.float32),
                   osy=2, osx=2, out0=out, macs_per_pair=hh * ww * 4 * c * c)
        return out

    # ---------------- encoder: both streams per launch; level 4 on image B only (SNUNet.py:119-129)
    hs = [h >> i for i in range(5)]
    ws = [w >> i for i in range(5)]
    x = {}                                    # node name -> tensor name
    pooled = p.tensor("p0", 2, hs[1], ws[1], f[0])
    x["0_0"] = nested("conv0_0", [L.Segment("in", in_ch)], f[0], hs[0], ws[0], pair=True, pool=pooled)
    for lvl in (1, 2, 3):
        nxt = p.tensor(f"p{lvl}", 2, hs[lvl + 1], ws[lvl + 1], f[lvl])
        x[f"{lvl}_0"] = nested(f"conv{lvl}_0", [L.Segment(pooled, f[lvl - 1])], f[lvl], hs[lvl], ws[lvl], pair=True, pool=nxt)
        pooled = nxt
    x["4_0"] = nested("conv4_0", [L.Segment(pooled, f[3], stream=1)], f[4], hs[4], ws[4], pair=False)

    def both(node: str, lvl: int) -> List[L.Segment]:
        return [L.Segment(x[node], f[lvl], stream=0), L.Segment(x[node], f[lvl], stream=1)]

    def one(node: str, lvl: int) -> L.Segment:
        return L.Segment(x[node], f[lvl])

    def dec(node: str, lvl: int, segs: List[L.Segment]) -> None:
        x[node] = nested(f"conv{node}", segs, f[lvl], hs[lvl], ws[lvl], pair=False)

    def upnode(name: str, node: str, lvl: int, stream: int = 0) -> L.Segment:
        return L.Segment(upsample(name, x[node], f[lvl], hs[lvl], ws[lvl], stream), f[lvl])

    # ---------------- nested decoder, in the reference's order (SNUNet.py:131-142)
    dec("0_1", 0, both("0_0", 0) + [upnode("Up1_0", "1_0", 1, stream=1)])
    dec("1_1", 1, both("1_0", 1) + [upnode("Up2_0", "2_0", 2, stream=1)])
    dec("0_2", 0, both("0_0", 0) + [one("0_1", 0), upnode("Up1_1", "1_1", 1)])
    dec("2_1", 2, both("2_0", 2) + [upnode("Up3_0", "3_0", 3, stream=1)])
    dec("1_2", 1, both("1_0", 1) + [one("1_1", 1), upnode("Up2_1", "2_1", 2)])
    dec("0_3", 0, both("0_0", 0) + [one("0_1", 0), one("0_2", 0), upnode("Up1_2", "1_2", 1)])
    dec("3_1", 3, both("3_0", 3) + [upnode("Up4_0", "4_0", 4)])
    dec("2_2", 2, both("2_0", 2) + [one("2_1", 2), upnode("Up3_1", "3_1", 3)])
    dec("1_3", 1, both("1_0", 1) + [one("1_1", 1), one("1_2", 1), upnode("Up2_2", "2_2", 2)])
    dec("0_4", 0, both("0_0", 0) + [one("0_1", 0), one("0_2", 0), one("0_3", 0), upnode("Up1_3", "1_3", 1)])

    # ---------------- ECAM + conv_final (SNUNet.py:144-149)
    c4 = 4 * f[0]
    p.ops.append(L.EcamHeadSpec(
        name="ecam_head", srcs=[x["0_1"], x["0_2"], x["0_3"], x["0_4"]], c=f[0], n_class=out_ch,
        ca_fc1=sd["ca.fc1.weight"].reshape(-1, c4).numpy().astype(np.float32).copy(),
        ca_fc2=sd["ca.fc2.weight"].reshape(c4, -1).numpy().astype(np.float32).copy(),
        ca1_fc1=sd["ca1.fc1.weight"].reshape(-1, f[0]).numpy().astype(np.float32).copy(),
        ca1_fc2=sd["ca1.fc2.weight"].reshape(f[0], -1).numpy().astype(np.float32).copy(),
        w_final=sd["conv_final.weight"].reshape(out_ch, c4).numpy().astype(np.float32).copy(),
        b_final=sd["conv_final.bias"].numpy().astype(np.float32).copy(),
        out_ext=0, macs_per_pair=h * w * c4 * out_ch, split=p.split))
    p.ext.append(L.ExtOutput("logits", out_ch, h, w))
    return p
