"""ChangeFormerV6 (MiT-style Siamese transformer encoder + multi-scale difference decoder) behind ``net_G(x1, x2)``.

Drop-in for ``models/ChangeFormer.py::ChangeFormerV6`` (registry key ``ChangeFormerV6``, models/networks.py:190-191):
same constructor arguments, the reference's parameter names (``Tenc_x2.block1.0.attn.q.weight`` ...
``TDec_x2.linear_c4.proj.weight`` ...: a reference ``state_dict`` loads), same return value: a list of five tensors
with the full-resolution logits last.  Eval mode (Dropout / DropPath are identities).

Lowering (both temporal images ride through every encoder launch):

* tokens ``[B, N, C]`` are pixels ``[B, H, W, C]``: every ``nn.Linear`` is a 1x1 conv of the implicit-GEMM kernel with the
  bias, GELU and the residual add in its epilogue;
* OverlapPatchEmbed (ChangeFormer.py:195-236): the 7x7 stride-4 conv on the image reads strided TMA boxes (one per tap),
  the 7x7 stride-2 convs read the previous stage's output stored space-to-depth (4x4 taps over parity classes);
* LayerNorm is one bandwidth kernel (fp32 statistics), optionally writing the space-to-depth copy as well;
* spatial-reduction attention (:298-358): ``sr`` is a stride-``sr`` conv (strided boxes), ``kv`` a 1x1 conv on the 64
  reduced tokens, and ``softmax(q k^T / sqrt(d)) v`` one kernel that keeps the 64 keys/values of a head in shared memory
  (the score matrix never exists in HBM);
* ``Mlp`` (:260-295): fc1 -> depth-wise 3x3 + bias + GELU (one bandwidth kernel) -> fc2 + residual;
* the decoder is ChangeGNNV1's (``changevig.lower_diff_decoder``).
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np
import torch
import torch.nn as nn

from . import lowering as L
from .changevig import _DecoderV1, lower_diff_decoder
from .module import PlannedModule

_DIMS = (64, 128, 320, 512)
_DEPTHS = (3, 3, 4, 3)
_HEADS = (1, 2, 4, 8)
_SRS = (8, 4, 2, 1)


class _OverlapPatchEmbed(nn.Module):
    def __init__(self, k: int, stride: int, cin: int, cout: int):
        super().__init__()
        self.proj = nn.Conv2d(cin, cout, kernel_size=k, stride=stride, padding=k // 2)
        self.norm = nn.LayerNorm(cout)


class _Attention(nn.Module):
    def __init__(self, dim: int, sr: int):
        super().__init__()
        self.q = nn.Linear(dim, dim, bias=True)
        self.kv = nn.Linear(dim, dim * 2, bias=True)
        self.proj = nn.Linear(dim, dim)
        if sr > 1:
            self.sr = nn.Conv2d(dim, dim, kernel_size=sr, stride=sr)
            self.norm = nn.LayerNorm(dim)


class _DWConv(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        self.dwconv = nn.Conv2d(dim, dim, 3, 1, 1, bias=True, groups=dim)


class _Mlp(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        self.fc1 = nn.Linear(dim, 4 * dim)
        self.dwconv = _DWConv(4 * dim)
        self.fc2 = nn.Linear(4 * dim, dim)


class _Block(nn.Module):
    def __init__(self, dim: int, sr: int):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = _Attention(dim, sr)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = _Mlp(dim)


class _EncoderTransformerV3(nn.Module):
    """models/ChangeFormer.py:1340-1432 (parameters only)."""

    def __init__(self, in_chans: int, depths=_DEPTHS, patch_k: int = 7, intra_patch: bool = False):
        super().__init__()
        self.patch_embed1 = _OverlapPatchEmbed(7, 4, in_chans, _DIMS[0])
        self.patch_embed2 = _OverlapPatchEmbed(patch_k, 2, _DIMS[0], _DIMS[1])
        self.patch_embed3 = _OverlapPatchEmbed(patch_k, 2, _DIMS[1], _DIMS[2])
        self.patch_embed4 = _OverlapPatchEmbed(patch_k, 2, _DIMS[2], _DIMS[3])
        for s in range(4):
            setattr(self, f"block{s + 1}", nn.ModuleList([_Block(_DIMS[s], _SRS[s]) for _ in range(depths[s])]))
            setattr(self, f"norm{s + 1}", nn.LayerNorm(_DIMS[s], eps=1e-6))
            if intra_patch and s < 3:
                # EncoderTransformer's "intra-patch encoder" blocks (ChangeFormer.py:51-58, 66-73, 81-88): constructed, never
                # called by forward_features (:141-185) -- parameters only
                setattr(self, f"patch_block{s + 1}", nn.ModuleList([_Block(_DIMS[s + 1], _SRS[s])]))
                setattr(self, f"pnorm{s + 1}", nn.LayerNorm(_DIMS[s + 1], eps=1e-6))


class ChangeFormerV6(PlannedModule):
    """models/ChangeFormer.py:1669-1701."""
    default_chunk_pairs = 32

    def __init__(self, input_nc: int = 3, output_nc: int = 2, decoder_softmax: bool = False, embed_dim: int = 256):
        super().__init__()
        if decoder_softmax:
            raise NotImplementedError("stcd_b200.ChangeFormerV6 serves decoder_softmax=False (networks.py:191)")
        if input_nc > 8 or output_nc > 8 or embed_dim % 16:
            raise NotImplementedError("input_nc, output_nc <= 8 and embed_dim a multiple of 16")
        self.input_nc, self.output_nc = input_nc, output_nc
        self.embed_dims = list(_DIMS)
        self.depths = list(_DEPTHS)
        self.embedding_dim = embed_dim
        self.Tenc_x2 = _EncoderTransformerV3(input_nc)
        self.TDec_x2 = _DecoderV1(_DIMS, embed_dim, output_nc, head="linear_c")
        for m in self.Tenc_x2.modules():                    # EncoderTransformer_v3._init_weights, :1404-1418
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=.02)
                nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.Conv2d):
                fan_out = m.kernel_size[0] * m.kernel_size[1] * m.out_channels // m.groups
                m.weight.data.normal_(0, (2.0 / fan_out) ** 0.5)
                m.bias.data.zero_()

    def lower(self, h: int, w: int) -> L.Program:
        return lower_changeformer(self.state_dict(), self.input_nc, self.embedding_dim, self.output_nc, h, w)

    @torch.no_grad()
    def forward(self, x1: torch.Tensor, x2: torch.Tensor):
        return self.plan_for(x1).forward(x1, x2)        # list of 5, full-resolution logits last

    def _wrap_outputs(self, outs):
        return list(outs)


# ------------------------------------------------------------------------------------------
def lower_changeformer(sd: Dict[str, torch.Tensor], in_ch: int, e: int, n_class: int, h: int, w: int) -> L.Program:
    """state_dict of the reference ChangeFormerV6 -> fused-op Program (eval mode)."""
    if h % 256 or w % 256:
        # stage 1 reduces its keys by sr = 8 at 1/4 scale, ..., and every stage must leave an integer, even token grid
        raise ValueError(f"ChangeFormerV6 lowering needs H and W divisible by 256 (got {h}x{w}); the reference is built for 256")
    if (h // 32) * (w // 32) > 64:
        raise ValueError("the attention kernel keeps at most 64 reduced tokens per image (256x256 inputs, the reference's img_size)")
    sd = {k: v.detach().to("cpu", torch.float32) for k, v in sd.items()}
    p = L.Program(model="ChangeFormerV6", in_channels=in_ch, h=h, w=w)
    feats = lower_mit_encoder(p, sd, "Tenc_x2", in_ch, _DEPTHS, h, w)
    lower_diff_decoder(p, sd, feats, e, n_class, "TDec_x2", "TDec_x2.linear_c{k}")
    return p


def lower_mit_encoder(p: L.Program, sd: Dict[str, torch.Tensor], enc: str, in_ch: int, depths, h: int, w: int):
    """EncoderTransformer (ChangeFormer.py:23-193; patch embeds 7/s4 then 3/s2) == EncoderTransformer_v3 (:1342-1472; 7/s4 then 7/s2):
    the patch size is read off the weights.  Returns [(tensor, channels, h, w)] per stage, both streams."""
    ones = lambda c: np.ones(c, np.float32)  # noqa: E731
    npf = lambda t: t.numpy().astype(np.float32)  # noqa: E731

    def linear(name: str, pre: str, src: str, cin: int, cout: int, hh: int, ww: int, **kw) -> None:
        wt = sd[f"{pre}.weight"][:, :, None, None]
        L.add_conv(p, name, [L.Segment(src, cin)], L.conv_taps(wt, pad=0), cout, hh, ww, 1, ones(cout), npf(sd[f"{pre}.bias"]),
                   pair=True, macs_per_pair=2 * hh * ww * cin * cout, **kw)

    def layernorm(name: str, pre: str, src: str, dst: str, c: int, eps: float, dst_s2d=None) -> None:
        p.ops.append(L.LayerNormSpec(name, src, dst, c, npf(sd[f"{pre}.weight"]), npf(sd[f"{pre}.bias"]), eps, dst_s2d))

    p.tensor("in", 2, h, w, 8)
    p.ops.append(L.InputPackSpec("pack", "in", in_ch))
    feats = []
    x_s2d = None
    hh, ww = h, w
    for s in range(4):
        c, heads, sr = _DIMS[s], _HEADS[s], _SRS[s]
        pe = f"{enc}.patch_embed{s + 1}"
        wt = sd[f"{pe}.proj.weight"]
        if s == 0:
            hh, ww = hh // 4, ww // 4
            t = p.tensor(f"{pe}.conv", 2, hh, ww, c)
            L.add_conv(p, f"{pe}.proj", [L.Segment("in", in_ch, sy=4, sx=4)], L.conv_taps(wt, pad=3), c, hh, ww, 1, ones(c),
                       npf(sd[f"{pe}.proj.bias"]), pair=True, out0=t, macs_per_pair=2 * hh * ww * 49 * in_ch * c)
        else:
            cprev = _DIMS[s - 1]
            hh, ww = hh // 2, ww // 2
            t = p.tensor(f"{pe}.conv", 2, hh, ww, c)
            kk = wt.shape[2]
            L.add_conv(p, f"{pe}.proj", L.s2d_segments(x_s2d, cprev), [(0, 0, L.s2d_conv_taps(wt, pad=kk // 2))], c, hh, ww, 1, ones(c),
                       npf(sd[f"{pe}.proj.bias"]), pair=True, out0=t, macs_per_pair=2 * hh * ww * kk * kk * cprev * c)
        x = p.tensor(f"{pe}.out", 2, hh, ww, c)
        layernorm(f"{pe}.norm", f"{pe}.norm", t, x, c, 1e-5)
        for i in range(depths[s]):
            blk = f"{enc}.block{s + 1}.{i}"
            a = f"{blk}.attn"
            n1 = p.tensor(f"{blk}.n1", 2, hh, ww, c)
            layernorm(f"{blk}.norm1", f"{blk}.norm1", x, n1, c, 1e-6)
            q = p.tensor(f"{a}.q", 2, hh, ww, c)
            linear(f"{a}.q", f"{a}.q", n1, c, c, hh, ww, out0=q)
            if sr > 1:
                hk, wk = hh // sr, ww // sr
                srt = p.tensor(f"{a}.sr", 2, hk, wk, c)
                L.add_conv(p, f"{a}.sr", [L.Segment(n1, c, sy=sr, sx=sr)], L.conv_taps(sd[f"{a}.sr.weight"], pad=0), c, hk, wk, 1,
                           ones(c), npf(sd[f"{a}.sr.bias"]), pair=True, out0=srt, macs_per_pair=2 * hk * wk * sr * sr * c * c)
                kin = p.tensor(f"{a}.srn", 2, hk, wk, c)
                layernorm(f"{a}.norm", f"{a}.norm", srt, kin, c, 1e-5)
            else:
                hk, wk, kin = hh, ww, n1
            kv = p.tensor(f"{a}.kv", 2, hk, wk, 2 * c)
            linear(f"{a}.kv", f"{a}.kv", kin, c, 2 * c, hk, wk, out0=kv)
            ao = p.tensor(f"{a}.o", 2, hh, ww, c)
            d = c // heads
            p.ops.append(L.AttentionSpec(f"{a}.softmax", q, kv, ao, c, heads, float(d ** -0.5),
                                         macs_per_pair=2 * 2 * hh * ww * hk * wk * c))
            x1 = p.tensor(f"{blk}.x1", 2, hh, ww, c)
            linear(f"{a}.proj", f"{a}.proj", ao, c, c, hh, ww, res=x, out0=x1)
            n2 = p.tensor(f"{blk}.n2", 2, hh, ww, c)
            layernorm(f"{blk}.norm2", f"{blk}.norm2", x1, n2, c, 1e-6)
            m = f"{blk}.mlp"
            h1 = p.tensor(f"{m}.h1", 2, hh, ww, 4 * c)
            linear(f"{m}.fc1", f"{m}.fc1", n2, c, 4 * c, hh, ww, out0=h1)
            h2 = p.tensor(f"{m}.h2", 2, hh, ww, 4 * c)
            p.ops.append(L.DWConvSpec(f"{m}.dwconv", h1, h2, 4 * c, npf(sd[f"{m}.dwconv.dwconv.weight"].reshape(4 * c, 9)),
                                      npf(sd[f"{m}.dwconv.dwconv.bias"]), gelu=True, macs_per_pair=2 * hh * ww * 9 * 4 * c))
            x = p.tensor(f"{blk}.out", 2, hh, ww, c)
            linear(f"{m}.fc2", f"{m}.fc2", h2, 4 * c, c, hh, ww, res=x1, out0=x)
        f = p.tensor(f"{enc}.f{s + 1}", 2, hh, ww, c)
        x_s2d = p.tensor(f"{enc}.f{s + 1}_s2d", 2, hh // 2, ww // 2, 4 * c) if s < 3 else None
        layernorm(f"{enc}.norm{s + 1}", f"{enc}.norm{s + 1}", x, f, c, 1e-6, dst_s2d=x_s2d)
        feats.append((f, c, hh, ww))
    return feats
