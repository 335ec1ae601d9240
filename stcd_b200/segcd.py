"""smp ``SegCD`` (Siamese Unet over a ResNet encoder) behind the reference's ``net(A, B)`` contract.

Drop-in for ``segmentation_models_pytorch.SegCD`` (decoders/unet/model.py:267-332), the model the STCD
scripts instantiate (train_stcd.py:637-638): same constructor keywords, same parameter names
(``encoder.layer1.0.conv1.weight`` ... ``decoder.blocks.0.conv1.0.weight`` ...
``segmentation_head.0.weight``: a reference ``state_dict`` loads), same return value
``(mask_t1, mask_t2, change)``.  Eval-mode only; encoders resnet18 / resnet34 (BasicBlock) and resnet50
(Bottleneck: the encoder train_stcd.py:638 actually selects).

Lowering (both temporal images ride through every launch as Siamese pair tiles sharing the weights):

* the 7x7 stride-2 stem (torchvision ResNet.conv1; smp/encoders/resnet.py:50) reads a
  space-to-depth packing of the input (12 channels at half resolution): it becomes a 4x4 stride-1
  conv with halo reuse instead of 49 strided taps;
* feature maps that are only read by stride-2 convs and by the decoder (the stem output and the last
  block of layer1..3) are STORED space-to-depth by the producing conv's epilogue, so stride-2 3x3 /
  1x1 convs are stride-1 convs over parity classes (9 / 1 taps), never strided loads;
* ``F.interpolate(x, 2, "nearest")`` + ``torch.cat([x, skip])`` + conv (decoders/unet/decoder.py:35-40)
  is never materialised: the conv runs as 4 output phases, the up-sampled half with merged 2x2 taps
  on the low-resolution tensor (4/9 of the MACs), the skip half on the parity classes;
* BatchNorm folds into the epilogue, the BasicBlock residual add + ReLU too (models/resnet.py:57-75);
* the three ``segmentation_head`` calls and ``min(head(|d1-d2|), |m1-m2|)`` (model.py:321-330) are one
  bandwidth kernel.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn as nn

from . import lowering as L
from .module import PlannedModule

_LAYERS = {"resnet18": (2, 2, 2, 2), "resnet34": (3, 4, 6, 3), "resnet50": (3, 4, 6, 3)}
_BOTTLENECK = {"resnet50"}           # torchvision Bottleneck (expansion 4); the others use BasicBlock
_WIDTHS = (64, 128, 256, 512)


class _BasicBlock(nn.Module):
    """Parameter holder with torchvision BasicBlock's names (≡ models/resnet.py:37-75)."""

    def __init__(self, cin: int, cout: int, stride: int):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride=stride, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(cout)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(cout)
        self.downsample = None
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride=stride, bias=False), nn.BatchNorm2d(cout))


class _Bottleneck(nn.Module):
    """Parameter holder with torchvision Bottleneck's names (== models/resnet.py:78-124; stride on conv2)."""
    expansion = 4

    def __init__(self, cin: int, width: int, stride: int):
        super().__init__()
        cout = width * self.expansion
        self.conv1 = nn.Conv2d(cin, width, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(width)
        self.conv2 = nn.Conv2d(width, width, 3, stride=stride, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(width)
        self.conv3 = nn.Conv2d(width, cout, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(cout)
        self.downsample = None
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride=stride, bias=False), nn.BatchNorm2d(cout))


class _ResNetEncoder(nn.Module):
    """smp/encoders/resnet.py:37-65 (parameters only; fc / avgpool are deleted upstream too)."""

    def __init__(self, name: str, in_channels: int):
        super().__init__()
        self.conv1 = nn.Conv2d(in_channels, 64, 7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        cin = 64
        bott = name in _BOTTLENECK
        for li, (n, width) in enumerate(zip(_LAYERS[name], _WIDTHS)):
            blocks = []
            for b in range(n):
                stride = 2 if (b == 0 and li > 0) else 1
                blocks.append(_Bottleneck(cin, width, stride) if bott else _BasicBlock(cin, width, stride))
                cin = width * 4 if bott else width
            setattr(self, f"layer{li + 1}", nn.Sequential(*blocks))
        e = 4 if bott else 1
        self.out_channels = (in_channels, 64, 64 * e, 128 * e, 256 * e, 512 * e)


def _conv_bn_relu(cin: int, cout: int) -> nn.Sequential:
    """smp Conv2dReLU with use_batchnorm=True (base/modules.py:10-47): conv (no bias), BN, ReLU."""
    return nn.Sequential(nn.Conv2d(cin, cout, 3, padding=1, bias=False), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))


class _DecoderBlock(nn.Module):
    def __init__(self, cin: int, cskip: int, cout: int):
        super().__init__()
        self.conv1 = _conv_bn_relu(cin + cskip, cout)
        self.conv2 = _conv_bn_relu(cout, cout)


class _UnetDecoder(nn.Module):
    """decoders/unet/decoder.py:67-123 (center = Identity for ResNet encoders)."""

    def __init__(self, encoder_channels: Sequence[int], decoder_channels: Sequence[int]):
        super().__init__()
        enc = list(encoder_channels[1:])[::-1]
        cin = [enc[0]] + list(decoder_channels[:-1])
        cskip = enc[1:] + [0]
        self.blocks = nn.ModuleList([_DecoderBlock(a, b, c) for a, b, c in zip(cin, cskip, decoder_channels)])


class SegCD(PlannedModule):
    """segmentation_models_pytorch.SegCD, decoders/unet/model.py:267-332."""
    default_chunk_pairs = 2

    def __init__(self, encoder_name: str = "resnet34", encoder_depth: int = 5, encoder_weights: Optional[str] = None,
                 decoder_use_batchnorm: bool = True, decoder_channels: Sequence[int] = (256, 128, 64, 32, 16),
                 decoder_attention_type: Optional[str] = None, in_channels: int = 3, classes: int = 1,
                 activation=None, aux_params: Optional[dict] = None):
        super().__init__()
        if encoder_name not in _LAYERS:
            raise NotImplementedError(f"stcd_b200.SegCD serves the ResNet encoders {sorted(_LAYERS)}; "
                                      f"'{encoder_name}' stays with the reference")
        if encoder_weights is not None:
            raise NotImplementedError("pretrained encoder weights need the network; load a state_dict instead")
        if encoder_depth != 5 or len(decoder_channels) != 5:
            raise NotImplementedError("encoder_depth must be 5 (the reference's default)")
        if decoder_use_batchnorm is not True or decoder_attention_type is not None or activation is not None or aux_params:
            raise NotImplementedError("stcd_b200.SegCD serves decoder_use_batchnorm=True, no attention, no activation, no aux head")
        if classes != 1:
            raise NotImplementedError("stcd_b200.SegCD serves classes=1 (the STCD scripts' setting, train_stcd.py:637)")
        if in_channels > 4:
            raise NotImplementedError("in_channels <= 4")
        self.encoder_name = encoder_name
        self.inchannels = in_channels
        self.decoder_channels = tuple(decoder_channels)
        self.encoder = _ResNetEncoder(encoder_name, in_channels)
        self.encoder_channels = self.encoder.out_channels
        self.decoder = _UnetDecoder(self.encoder_channels, decoder_channels)
        self.segmentation_head = nn.Sequential(nn.Conv2d(decoder_channels[-1], classes, 3, padding=1), nn.Identity(),
                                               nn.Identity())
        self.name = "u-{}".format(encoder_name)
        self.initialize()

    def initialize(self) -> None:
        """base/initialization.py:4-27 on the decoder and head (the encoder keeps torchvision's init)."""
        for m in self.encoder.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
        for m in self.decoder.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_uniform_(m.weight, mode="fan_in", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        for m in self.segmentation_head.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.xavier_uniform_(m.weight)
                nn.init.constant_(m.bias, 0)

    ffctl = False        # FFCTLCD: feature-level branch = head(decoder(|f1 - f2|)) instead of head(|d1 - d2|)

    def lower(self, h: int, w: int) -> L.Program:
        return lower_segcd(self.state_dict(), self.encoder_name, self.inchannels, self.decoder_channels, h, w, ffctl=self.ffctl)

    @torch.no_grad()
    def forward(self, A: torch.Tensor, B: torch.Tensor):
        m1, m2, change = self.plan_for(A).forward(A, B)
        return m1, m2, change


class FFCTLCD(SegCD):
    """segmentation_models_pytorch.FFCTLCD, decoders/unet/model.py:335-423 (the alternative train_stcd.py:636 keeps
    commented out): same parameters as SegCD; the feature-level change branch runs the decoder on |f1 - f2| of every
    encoder feature, the decision-level branch is |mask_t1 - mask_t2|, change = min of the two."""
    ffctl = True


# ------------------------------------------------------------------------------------------
def stem_s2d_taps(weight: torch.Tensor) -> List:
    """7x7 stride-2 pad-3 conv over x  ==  4x4 stride-1 conv over space-to-depth(x):
    out(i, j) = sum W[c, ky, kx] x(c, 2i + ky - 3, 2j + kx - 3); with ky - 3 = 2 dy + py the tap
    (dy, dx) in [-2, 1]^2 reads channel (py*2 + px)*cin + c of s2d(x) at (i + dy, j + dx)."""
    cout, cin, k, _ = weight.shape
    pad = k // 2
    taps = {}
    for ky in range(k):
        dy, py = divmod(ky - pad, 2)
        for kx in range(k):
            dx, px = divmod(kx - pad, 2)
            wt = taps.setdefault((dy, dx), torch.zeros(cout, 4 * cin, dtype=torch.float32))
            wt[:, (py * 2 + px) * cin: (py * 2 + px + 1) * cin] = weight[:, :, ky, kx].to(torch.float32)
    return [(0, 0, [(dy, dx, wt) for (dy, dx), wt in sorted(taps.items())])]


def lower_segcd(sd: Dict[str, torch.Tensor], encoder_name: str, in_channels: int, decoder_channels: Sequence[int],
                h: int, w: int, ffctl: bool = False) -> L.Program:
    """state_dict of the reference SegCD -> fused-op Program (eval mode)."""
    if h % 32 or w % 32:
        raise ValueError(f"SegCD lowering needs H and W divisible by 32 (got {h}x{w}): the reference's decoder "
                         "cannot concatenate its skips otherwise")
    sd = {k: v.detach().to("cpu", torch.float32) for k, v in sd.items()}
    p = L.Program(model=f"SegCD-{encoder_name}", in_channels=in_channels, h=h, w=w)
    ones = lambda c: np.ones(c, np.float32)  # noqa: E731

    def bn(prefix: str, c: int):
        return L.fold_bn(None, L.bn_params(sd, prefix), c)

    # ---------------- stem: space-to-depth pack, 4x4 conv, BN, ReLU -> f1 stored space-to-depth
    hh, ww = h // 2, w // 2
    p.tensor("in", 2, hh, ww, 16)
    p.ops.append(L.InputPackSpec("pack", "in", in_channels, s2d=True))
    f1s = p.tensor("f1s", 2, hh // 2, ww // 2, 4 * 64)
    sc, sh = bn("encoder.bn1", 64)
    wt = sd["encoder.conv1.weight"]
    L.add_conv(p, "encoder.conv1", [L.Segment("in", 4 * in_channels)], stem_s2d_taps(wt), 64, hh, ww, 1, sc, sh,
               pair=True, relu=True, out0=f1s, out0_s2d=True, macs_per_pair=2 * hh * ww * 64 * in_channels * 49)
    skips = [(f1s, 64)]                      # (space-to-depth tensor, channels), shallow -> deep
    hh, ww = hh // 2, ww // 2
    x = p.tensor("p1", 2, hh, ww, 64)
    p.ops.append(L.MaxPoolS2DSpec("encoder.maxpool", f1s, x, 64))
    x_s2d = False
    cin = 64

    # ---------------- residual layers
    bott = encoder_name in _BOTTLENECK
    for li, (n_blocks, width) in enumerate(zip(_LAYERS[encoder_name], _WIDTHS)):
        cout = width * 4 if bott else width
        for b in range(n_blocks):
            pre = f"encoder.layer{li + 1}.{b}"
            stride = 2 if (b == 0 and li > 0) else 1
            last_of_layer = (b == n_blocks - 1) and li < 3
            h_in, w_in = hh, ww
            if stride == 2:
                assert x_s2d
                hh, ww = hh // 2, ww // 2
            else:
                assert not x_s2d
            if bott:
                w1, w2, w3 = sd[f"{pre}.conv1.weight"], sd[f"{pre}.conv2.weight"], sd[f"{pre}.conv3.weight"]
                sc, sh = bn(f"{pre}.bn1", width)
                t2 = p.tensor(f"{pre}.t2", 2, hh, ww, width)
                if stride == 2:
                    # 1x1 conv at the input resolution on a space-to-depth input: one launch per parity class,
                    # written to the same class of a space-to-depth output, which the stride-2 3x3 then reads
                    t1 = p.tensor(f"{pre}.t1_s2d", 2, hh, ww, 4 * width)
                    for cls, seg in enumerate(L.s2d_segments(x, cin)):
                        L.add_conv(p, f"{pre}.conv1.{cls}", [seg], L.conv_taps(w1, pad=0), width, hh, ww, 1, sc, sh, pair=True,
                                   relu=True, out0=t1, out0_coff=cls * width, macs_per_pair=2 * hh * ww * cin * width)
                    sc, sh = bn(f"{pre}.bn2", width)
                    L.add_conv(p, f"{pre}.conv2", L.s2d_segments(t1, width), [(0, 0, L.s2d_conv_taps(w2, pad=1))], width, hh, ww, 1,
                               sc, sh, pair=True, relu=True, out0=t2, macs_per_pair=2 * hh * ww * 9 * width * width)
                else:
                    t1 = p.tensor(f"{pre}.t1", 2, hh, ww, width)
                    L.add_conv(p, f"{pre}.conv1", [L.Segment(x, cin)], L.conv_taps(w1, pad=0), width, hh, ww, 1, sc, sh, pair=True,
                               relu=True, out0=t1, macs_per_pair=2 * hh * ww * cin * width)
                    sc, sh = bn(f"{pre}.bn2", width)
                    L.add_conv(p, f"{pre}.conv2", [L.Segment(t1, width)], L.conv_taps(w2, pad=1), width, hh, ww, 1, sc, sh,
                               pair=True, relu=True, out0=t2, macs_per_pair=2 * hh * ww * 9 * width * width)
                last_w, last_in, last_c, last_bn, last_name = w3, t2, width, f"{pre}.bn3", f"{pre}.conv3"
                last_taps, last_macs = L.conv_taps(w3, pad=0), 2 * hh * ww * width * cout
            else:
                w1, w2 = sd[f"{pre}.conv1.weight"], sd[f"{pre}.conv2.weight"]
                t = p.tensor(f"{pre}.t", 2, hh, ww, cout)
                sc, sh = bn(f"{pre}.bn1", cout)
                if stride == 2:
                    L.add_conv(p, f"{pre}.conv1", L.s2d_segments(x, cin), [(0, 0, L.s2d_conv_taps(w1, pad=1))], cout, hh, ww, 1,
                               sc, sh, pair=True, relu=True, out0=t, macs_per_pair=2 * hh * ww * 9 * cin * cout)
                else:
                    L.add_conv(p, f"{pre}.conv1", [L.Segment(x, cin)], L.conv_taps(w1, pad=1), cout, hh, ww, 1, sc, sh,
                               pair=True, relu=True, out0=t, macs_per_pair=2 * hh * ww * 9 * cin * cout)
                last_in, last_c, last_bn, last_name = t, cout, f"{pre}.bn2", f"{pre}.conv2"
                last_taps, last_macs = L.conv_taps(w2, pad=1), 2 * hh * ww * 9 * cout * cout
            # identity path: the input, or a 1x1 (stride s) conv + BN of it
            if f"{pre}.downsample.0.weight" in sd:
                ident = p.tensor(f"{pre}.ds", 2, hh, ww, cout)
                sc, sh = bn(f"{pre}.downsample.1", cout)
                wd = sd[f"{pre}.downsample.0.weight"]
                if stride == 2:
                    L.add_conv(p, f"{pre}.downsample", L.s2d_segments(x, cin), [(0, 0, L.s2d_conv_taps(wd, pad=0))], cout, hh, ww,
                               1, sc, sh, pair=True, out0=ident, macs_per_pair=2 * hh * ww * cin * cout)
                else:
                    L.add_conv(p, f"{pre}.downsample", [L.Segment(x, cin)], L.conv_taps(wd, pad=0), cout, hh, ww, 1, sc, sh,
                               pair=True, out0=ident, macs_per_pair=2 * hh * ww * cin * cout)
            else:
                ident = x
            sc, sh = bn(last_bn, cout)
            if last_of_layer:
                o = p.tensor(f"{pre}.o_s2d", 2, hh // 2, ww // 2, 4 * cout)
                skips.append((o, cout))
            else:
                o = p.tensor(f"{pre}.o", 2, hh, ww, cout)
            L.add_conv(p, last_name, [L.Segment(last_in, last_c)], last_taps, cout, hh, ww, 1, sc, sh, pair=True, relu=True,
                       res=ident, out0=o, out0_s2d=last_of_layer, macs_per_pair=last_macs)
            x, x_s2d, cin = o, last_of_layer, cout

    # ---------------- Unet decoder: (nearest x2, cat skip, conv-BN-ReLU, conv-BN-ReLU) x 5
    def decode(x, cin, hh, ww, skips, pair: bool, tag: str):
        """One pass of the decoder over (x, skips): both temporal streams as pair tiles, or one single-stream tensor set."""
        mult = 2 if pair else 1
        for bi, cout in enumerate(decoder_channels):
            pre = f"decoder.blocks.{bi}"
            skip = skips[len(skips) - 1 - bi] if bi < len(skips) else None
            w1 = sd[f"{pre}.conv1.0.weight"]
            cskip = skip[1] if skip else 0
            if w1.shape[1] != cin + cskip:
                raise ValueError(f"{pre}.conv1 expects {w1.shape[1]} input channels, lowering has {cin}+{cskip}")
            segs = [L.Segment(x, cin)] + (L.s2d_segments(skip[0], cskip) if skip else [])
            phases = []
            for a in range(2):
                for b in range(2):
                    st = L.SegTaps([L.up2_conv_taps(w1[:, :cin], 1, a, b)])
                    if skip:
                        st.extend(L.s2d_conv_taps(w1[:, cin:], 1, a, b))
                    phases.append((a, b, st))
            sc, sh = bn(f"{pre}.conv1.1", cout)
            cpad = (cout + 7) // 8 * 8
            t = p.tensor(f"{pre}.t{tag}", mult, 2 * hh, 2 * ww, cpad)
            L.add_conv(p, f"{pre}.conv1{tag}", segs, phases, cout, hh, ww, 1, sc, sh, pair=pair, relu=True, osy=2, osx=2, out0=t,
                       macs_per_pair=mult * 4 * hh * ww * 9 * (cin + cskip) * cout)
            hh, ww = 2 * hh, 2 * ww
            sc, sh = bn(f"{pre}.conv2.1", cout)
            o = p.tensor(f"{pre}.o{tag}", mult, hh, ww, cpad)
            L.add_conv(p, f"{pre}.conv2{tag}", [L.Segment(t, cout)], L.conv_taps(sd[f"{pre}.conv2.0.weight"], pad=1), cout, hh, ww, 1,
                       sc, sh, pair=pair, relu=True, out0=o, macs_per_pair=mult * hh * ww * 9 * cout * cout)
            x, cin = o, cout
        return x, cin, hh, ww

    diff_out = None
    if ffctl:
        # FFCTLCD (model.py:407-423): a third decoder pass over |f1 - f2| of every encoder feature
        dskips = []
        for (t_, c_) in skips:
            ts = p.tensors[t_]
            dt = p.tensor(f"{t_}.absdiff", 1, ts.h, ts.w, ts.c)
            p.ops.append(L.AbsDiffSpec(f"{t_}.absdiff", t_, dt, ts.c))
            dskips.append((dt, c_))
        xd = p.tensor(f"{x}.absdiff", 1, hh, ww, cin)
        p.ops.append(L.AbsDiffSpec(f"{x}.absdiff", x, xd, cin))
        diff_out, _, _, _ = decode(xd, cin, hh, ww, dskips, False, ".diff")
    x, cin, hh, ww = decode(x, cin, hh, ww, skips, True, "")

    # ---------------- heads + decision-level fusion
    wh = sd["segmentation_head.0.weight"]                      # [1, c, 3, 3]
    if cin not in (8, 16):
        raise NotImplementedError(f"head kernel serves 8 or 16 decoder channels (got {cin})")
    p.ops.append(L.SegHeadSpec("segmentation_head", x, cin,
                               wh[0].permute(1, 2, 0).reshape(9, cin).contiguous().numpy().astype(np.float32),
                               float(sd["segmentation_head.0.bias"][0]), out_ext=0, macs_per_pair=3 * hh * ww * 9 * cin,
                               diff_src=diff_out))
    for nm in ("mask_t1", "mask_t2", "change"):
        p.ext.append(L.ExtOutput(nm, 1, hh, ww))
    return p
