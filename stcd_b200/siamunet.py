"""FC-Siam-diff / FC-Siam-conc behind the reference's ``net_G(x1, x2)`` contract.

Drop-in for ``models/SiamUnet_diff.py::SiamUnet_diff`` and ``models/SiamUnet_conc.py::SiamUnet_conc``
(registry keys ``SiamUnet_abs`` / ``SiamUnet_conc``, models/networks.py:148-153): same ctor
arguments, same parameter names (so a reference ``state_dict`` loads), same return type (one
``[B, label_nbr, H, W]`` float32 tensor).  The forward lowers the eval-mode network to 25 fused
launches of libstcd_b200 per chunk of image pairs:

* both temporal images go through every encoder conv in ONE launch (Siamese pair tiles share
  the weight tiles); folded BatchNorm + ReLU, the 2x2 max-pool and the ``|f1 - f2|`` skip are
  written by the conv's epilogue (SiamUnet_diff.py:99-143,150);
* ``torch.cat`` is never materialised: decoder convs read (up-conv, skip) as K-segments;
* ``ConvTranspose2d(k3, s1, p1)`` is a conv with flipped/transposed weights, the stride-2
  up-convs run as 4 output phases (SiamUnet_diff.py:52-90).

Eval-mode semantics only (BatchNorm running statistics, Dropout2d = identity): this is the
inference hot path; training stays with the reference.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import lowering as L
from .module import PlannedModule

# (name, cin, cout) of the encoder convs per stage; decoder: (name, cout) chains after each up-conv.
_ENC: List[List[Tuple[str, int, int]]] = [
    [("11", -1, 16), ("12", 16, 16)],
    [("21", 16, 32), ("22", 32, 32)],
    [("31", 32, 64), ("32", 64, 64), ("33", 64, 64)],
    [("41", 64, 128), ("42", 128, 128), ("43", 128, 128)],
]
_DEC: List[Tuple[str, int, List[Tuple[str, int]]]] = [
    ("4", 128, [("43d", 128), ("42d", 128), ("41d", 64)]),
    ("3", 64, [("33d", 64), ("32d", 64), ("31d", 32)]),
    ("2", 32, [("22d", 32), ("21d", 16)]),
    ("1", 16, [("12d", 16), ("11d", -1)]),
]


def _skip_mult(fusion: str) -> int:
    """Width of the skip half of every decoder concat, in units of the encoder stage width."""
    return {"diff": 1, "conc": 2, "sub": 1, "cross": 1, "ef": 1}[fusion]


class _CrossConc(nn.Module):
    """models/SiamUnet_crossconc.py:11-33: channel-interleave (a0, b0, a1, b1, ...) -> grouped 3x3 conv (2C -> C, one
    group per channel pair) + BN + ReLU -> 3x3 conv C -> C + BN -> ReLU."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.in_channels = in_channels
        self.diff = nn.Sequential(nn.Conv2d(in_channels, in_channels // 2, kernel_size=3, padding=1, stride=1, groups=in_channels // 2),
                                  nn.BatchNorm2d(in_channels // 2), nn.ReLU())
        self.conv_res = nn.Sequential(nn.Conv2d(in_channels // 2, out_channels, kernel_size=3, padding=1, stride=1),
                                      nn.BatchNorm2d(out_channels))
        self.act = nn.ReLU()


class _SiamUnet(PlannedModule):
    fusion = "diff"
    supports_precision_path = True

    def __init__(self, input_nbr: int, label_nbr: int):
        super().__init__()
        self.input_nbr = input_nbr
        self.label_nbr = label_nbr
        for stage in _ENC:
            for name, cin, cout in stage:
                cin = (2 * input_nbr if self.fusion == "ef" else input_nbr) if cin < 0 else cin
                setattr(self, f"conv{name}", nn.Conv2d(cin, cout, kernel_size=3, padding=1))
                setattr(self, f"bn{name}", nn.BatchNorm2d(cout))
        for lvl, cup, chain in _DEC:
            setattr(self, f"upconv{lvl}", nn.ConvTranspose2d(cup, cup, kernel_size=3, padding=1, stride=2,
                                                             output_padding=1))
            cin = cup * (1 + _skip_mult(self.fusion))
            for name, cout in chain:
                cout = label_nbr if cout < 0 else cout
                setattr(self, f"conv{name}", nn.ConvTranspose2d(cin, cout, kernel_size=3, padding=1))
                if name != "11d":
                    setattr(self, f"bn{name}", nn.BatchNorm2d(cout))
                cin = cout
        if self.fusion == "cross":              # registered last, like the reference (SiamUnet_crossconc.py:119-122)
            for i, c in enumerate((16, 32, 64, 128)):
                setattr(self, f"cross_conc{i + 1}", _CrossConc(2 * c, c))

    def lower(self, h: int, w: int) -> L.Program:
        return lower_siamunet(self.state_dict(), self.fusion, self.input_nbr, self.label_nbr, h, w, precision=self.plan_precision)

    returns_list = False     # SiamUnet_sub / SiamUnet_cross_conc return [x11d] (SiamUnet_sub.py:177-180)

    @torch.no_grad()
    def forward(self, x1: torch.Tensor, x2: torch.Tensor):
        y = self.plan_for(x1).forward(x1, x2)[0]
        return [y] if self.returns_list else y

    def _wrap_outputs(self, outs):
        return [outs[0]] if self.returns_list else outs[0]


class SiamUnet_diff(_SiamUnet):
    """models/SiamUnet_diff.py:10-181."""
    fusion = "diff"


class SiamUnet_conc(_SiamUnet):
    """models/SiamUnet_conc.py:10-183."""
    fusion = "conc"


class SiamUnet_sub(_SiamUnet):
    """models/SiamUnet_sub.py:10-180: FC-Siam-diff with the signed skip ``x_2 - x_1`` (:150-173); returns ``[x11d]``."""
    fusion = "sub"
    returns_list = True


class SiamUnet_cross_conc(_SiamUnet):
    """models/SiamUnet_crossconc.py:36-212: the skip is ``cross_conc(x_1, x_2)``; returns ``[x11d]``."""
    fusion = "cross"
    returns_list = True


class Unet(_SiamUnet):
    """models/Unet.py:10-158 (FC-EF): ``cat(x1, x2)`` through ONE encoder; a plain U-Net decoder."""
    fusion = "ef"


# ------------------------------------------------------------------------------------------
def lower_siamunet(sd: Dict[str, torch.Tensor], fusion: str, input_nbr: int, label_nbr: int, h: int, w: int,
                   precision: str = "bf16") -> L.Program:
    """state_dict of the reference module -> fused-op Program (eval mode)."""
    if h % 16 or w % 16:
        # ReplicationPad2d (SiamUnet_diff.py:149) is a no-op exactly when H and W are multiples of 16
        raise ValueError(f"SiamUnet lowering needs H and W divisible by 16 (got {h}x{w})")
    if input_nbr > 16:
        raise ValueError("input_nbr > 16 not supported by the input packer")
    sd = {k: v.detach().to("cpu", torch.float32) for k, v in sd.items()}
    p = L.Program(model=f"SiamUnet_{fusion}", in_channels=input_nbr, h=h, w=w, precision=precision)
    if p.split and input_nbr > 8:
        raise NotImplementedError("split precision packs at most 8 input channels")
    p.tensor("in", 2, h, w, 8 if input_nbr <= 8 else 16)
    p.ops.append(L.InputPackSpec("pack", "in", input_nbr, split=p.split))

    def bn_fold(name: str, cout: int):
        return L.fold_bn(sd[f"conv{name}.bias"], L.bn_params(sd, f"bn{name}"), cout)

    # ---------------- encoder: both streams per launch (pair tiles); FC-EF: one stream over cat(x1, x2)
    ef = fusion == "ef"
    mult = 1 if ef else 2
    first_segs = [L.Segment("in", input_nbr, stream=0), L.Segment("in", input_nbr, stream=1)] if ef else [L.Segment("in", input_nbr)]
    cur, cur_c = "in", (2 * input_nbr if ef else input_nbr)
    hh, ww = h, w
    skips: List[Tuple[str, int, int, int]] = []   # (tensor, channels, h, w)
    for si, stage in enumerate(_ENC):
        for li, (name, _, cout) in enumerate(stage):
            last = li == len(stage) - 1
            scale, shift = bn_fold(name, cout)
            wt = sd[f"conv{name}.weight"]
            out0 = out_pool = out_diff = None
            if not last:
                out0 = p.tensor(f"x{name}", mult, hh, ww, cout)
            else:
                out_pool = p.tensor(f"x{si + 1}p", mult, hh // 2, ww // 2, cout)
                if fusion == "diff":
                    out_diff = p.tensor(f"d{si + 1}", 1, hh, ww, cout)
                    skips.append((out_diff, cout, hh, ww))
                else:
                    out0 = p.tensor(f"x{name}", mult, hh, ww, cout)
                    skips.append((out0, cout, hh, ww))
            segs = first_segs if cur == "in" else [L.Segment(cur, cur_c)]
            L.add_conv(p, f"conv{name}", segs, L.conv_taps(wt, pad=1), cout, hh, ww, 1,
                       scale, shift, pair=not ef, relu=True, out0=out0, out_pool=out_pool, out_diff=out_diff,
                       macs_per_pair=mult * hh * ww * 9 * cur_c * cout)
            cur, cur_c = (out0 if not last else out_pool), cout
        hh, ww = hh // 2, ww // 2

    if fusion == "cross":
        # cross_conc (SiamUnet_crossconc.py:11-33) per skip level: the grouped conv over the interleaved (a_c, b_c) pairs is
        # a dense conv over the two stream segments with diagonal [C x C] weights per tap; then 3x3 C -> C + BN, ReLU
        crossed = []
        for i, (sk, c, sh_, sw_) in enumerate(skips):
            cc = f"cross_conc{i + 1}"
            wg = sd[f"{cc}.diff.0.weight"]                       # [C, 2, 3, 3]: group c sees (a_c, b_c)
            dense = torch.zeros(c, 2 * c, 3, 3)
            idx = torch.arange(c)
            dense[idx, idx] = wg[:, 0]                           # stream-0 segment: a_c -> output c
            dense[idx, c + idx] = wg[:, 1]                       # stream-1 segment: b_c -> output c
            sc1, sh1 = L.fold_bn(sd[f"{cc}.diff.0.bias"], L.bn_params(sd, f"{cc}.diff.1"), c)
            ta = p.tensor(f"{cc}.a", 1, sh_, sw_, c)
            L.add_conv(p, f"{cc}.diff", [L.Segment(sk, c, stream=0), L.Segment(sk, c, stream=1)], L.conv_taps(dense, pad=1), c, sh_, sw_, 1,
                       sc1, sh1, relu=True, out0=ta, macs_per_pair=sh_ * sw_ * 9 * 2 * c)
            sc2, sh2 = L.fold_bn(sd[f"{cc}.conv_res.0.bias"], L.bn_params(sd, f"{cc}.conv_res.1"), c)
            tb = p.tensor(f"{cc}.out", 1, sh_, sw_, c)
            L.add_conv(p, f"{cc}.conv_res", [L.Segment(ta, c)], L.conv_taps(sd[f"{cc}.conv_res.0.weight"], pad=1), c, sh_, sw_, 1, sc2, sh2,
                       relu=True, out0=tb, macs_per_pair=sh_ * sw_ * 9 * c * c)
            crossed.append((tb, c, sh_, sw_))
        skips = crossed

    # ---------------- decoder: one stream; bottleneck = image 2 only (SiamUnet_diff.py:143,148)
    cur_stream = 0 if ef else 1
    for (lvl, cup, chain), (skip, skip_c, sh, sw) in zip(_DEC, reversed(skips)):
        wt = sd[f"upconv{lvl}.weight"]      # [cin, cout, 3, 3]
        up = p.tensor(f"u{lvl}", 1, sh, sw, cup)
        L.add_conv(p, f"upconv{lvl}", [L.Segment(cur, cur_c, stream=cur_stream)],
                   L.convT_phase_taps(wt, stride=2, pad=1), cup, sh // 2, sw // 2, 1,
                   np.ones(cup, np.float32), sd[f"upconv{lvl}.bias"].numpy().astype(np.float32),
                   osy=2, osx=2, out0=up, macs_per_pair=(sh // 2) * (sw // 2) * 9 * cur_c * cup)
        cur_stream = 0
        if fusion in ("diff", "cross", "ef"):
            segs = [L.Segment(up, cup), L.Segment(skip, skip_c)]
        else:      # conc: cat(up, x_1, x_2); sub: cat(up, x_2 - x_1) read as (x_1, x_2) with weights (-W, +W)
            segs = [L.Segment(up, cup), L.Segment(skip, skip_c, stream=0), L.Segment(skip, skip_c, stream=1)]
        cur, cur_c = None, None
        for ci, (name, cout) in enumerate(chain):
            cout = label_nbr if cout < 0 else cout
            wt = L.convT_as_conv_weight(sd[f"conv{name}.weight"])
            if fusion == "sub" and ci == 0:
                wt = torch.cat([wt[:, :cup], -wt[:, cup:], wt[:, cup:]], dim=1)
            cin = sum(s.c_real for s in segs)
            ref_cin = cin - skip_c if (fusion == "sub" and ci == 0) else cin        # MACs of the reference's layer
            if name == "11d":
                L.add_conv(p, f"conv{name}", segs, L.conv_taps(wt, pad=1), cout, sh, sw, 1,
                           np.ones(cout, np.float32), sd[f"conv{name}.bias"].numpy().astype(np.float32),
                           out_ext=0, macs_per_pair=sh * sw * 9 * ref_cin * cout)
                p.ext.append(L.ExtOutput("logits", cout, sh, sw))
            else:
                scale, shift = bn_fold(name, cout)
                out = p.tensor(f"x{name}", 1, sh, sw, cout)
                L.add_conv(p, f"conv{name}", segs, L.conv_taps(wt, pad=1), cout, sh, sw, 1, scale, shift,
                           relu=True, out0=out, macs_per_pair=sh * sw * 9 * ref_cin * cout)
                cur, cur_c = out, cout
                segs = [L.Segment(out, cout)]
    return p
