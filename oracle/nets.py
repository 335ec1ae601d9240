"""Functional fp32 CPU restatements of the reference forwards (eval mode).  TEST INFRASTRUCTURE.

Each function takes the reference module's ``state_dict`` and the image pair and returns what
the reference's ``forward(x1, x2)`` returns.  They are written against the reference's forward
code (cited per function) and pinned by tests/test_oracle.py against fixtures generated from
the real reference (oracle/make_golden.py).
"""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


def _bn(sd: SD, name: str, x: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    return F.batch_norm(x, sd[f"{name}.running_mean"], sd[f"{name}.running_var"], sd[f"{name}.weight"],
                        sd[f"{name}.bias"], training=False, eps=eps)


# ------------------------------------------------------------------------------------------
# FC-Siam-diff / FC-Siam-conc
def _siam_encoder(sd: SD, x: torch.Tensor) -> List[torch.Tensor]:
    """models/SiamUnet_diff.py:99-119 (one stream): returns [x12, x22, x33, x43, x4p]."""
    def cbr(name, t):
        return F.relu(_bn(sd, f"bn{name}", F.conv2d(t, sd[f"conv{name}.weight"], sd[f"conv{name}.bias"], padding=1)))
    feats = []
    for names in (("11", "12"), ("21", "22"), ("31", "32", "33"), ("41", "42", "43")):
        for n in names:
            x = cbr(n, x)          # Dropout2d is identity in eval mode
        feats.append(x)
        x = F.max_pool2d(x, kernel_size=2, stride=2)
    feats.append(x)
    return feats


def _cross_conc(sd: SD, name: str, a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """cross_conc.forward, models/SiamUnet_crossconc.py:27-33."""
    n, c, h, w = a.shape
    z = torch.ones(n, 2 * c, h, w)
    z[:, 0::2] = a
    z[:, 1::2] = b
    z = F.relu(_bn(sd, f"{name}.diff.1", F.conv2d(z, sd[f"{name}.diff.0.weight"], sd[f"{name}.diff.0.bias"], padding=1, groups=c)))
    return F.relu(_bn(sd, f"{name}.conv_res.1", F.conv2d(z, sd[f"{name}.conv_res.0.weight"], sd[f"{name}.conv_res.0.bias"], padding=1)))


def siamunet_forward(sd: SD, x1: torch.Tensor, x2: torch.Tensor, fusion: str) -> torch.Tensor:
    """models/SiamUnet_diff.py:94-181 ('diff'), SiamUnet_conc.py:94-183 ('conc'), SiamUnet_sub.py:94-180 ('sub'),
    SiamUnet_crossconc.py:124-212 ('cross'), Unet.py:92-158 ('ef': one encoder over cat(x1, x2))."""
    if fusion == "ef":
        f2 = _siam_encoder(sd, torch.cat((x1, x2), 1))
        f1 = f2
    else:
        f1 = _siam_encoder(sd, x1)
        f2 = _siam_encoder(sd, x2)
    x = f2[4]                      # the decoder's bottleneck comes from image 2 only (:143,148)

    def dcbr(name, t):
        return F.relu(_bn(sd, f"bn{name}", F.conv_transpose2d(t, sd[f"conv{name}.weight"], sd[f"conv{name}.bias"],
                                                               padding=1)))
    for lvl, names in ((4, ("43d", "42d", "41d")), (3, ("33d", "32d", "31d")), (2, ("22d", "21d")), (1, ("12d",))):
        a, b = f1[lvl - 1], f2[lvl - 1]
        x = F.conv_transpose2d(x, sd[f"upconv{lvl}.weight"], sd[f"upconv{lvl}.bias"], stride=2, padding=1,
                               output_padding=1)
        x = F.pad(x, (0, a.size(3) - x.size(3), 0, a.size(2) - x.size(2)), mode="replicate")
        skip = {"diff": lambda: (torch.abs(a - b),), "conc": lambda: (a, b), "sub": lambda: (torch.sub(b, a),),
                "cross": lambda: (_cross_conc(sd, f"cross_conc{lvl}", a, b),), "ef": lambda: (a,)}[fusion]()
        x = torch.cat((x,) + skip, 1)
        for n in names:
            x = dcbr(n, x)
    return F.conv_transpose2d(x, sd["conv11d.weight"], sd["conv11d.bias"], padding=1)


# ------------------------------------------------------------------------------------------
# SNUNet-CD (ECAM)
def _nested(sd: SD, name: str, x: torch.Tensor) -> torch.Tensor:
    """conv_block_nested.forward, models/SNUNet.py:17-26: the residual is conv1's PRE-BN output."""
    ident = F.conv2d(x, sd[f"{name}.conv1.weight"], sd[f"{name}.conv1.bias"], padding=1)
    y = F.relu(_bn(sd, f"{name}.bn1", ident))
    y = _bn(sd, f"{name}.bn2", F.conv2d(y, sd[f"{name}.conv2.weight"], sd[f"{name}.conv2.bias"], padding=1))
    return F.relu(y + ident)


def _up(sd: SD, name: str, x: torch.Tensor) -> torch.Tensor:
    """up.forward, models/SNUNet.py:29-43: ConvTranspose2d(C, C, 2, stride=2)."""
    return F.conv_transpose2d(x, sd[f"{name}.up.weight"], sd[f"{name}.up.bias"], stride=2)


def _channel_attention(sd: SD, name: str, x: torch.Tensor) -> torch.Tensor:
    """ChannelAttention.forward, models/SNUNet.py:46-59 (1x1 convs without bias)."""
    def mlp(v):
        return F.conv2d(F.relu(F.conv2d(v, sd[f"{name}.fc1.weight"])), sd[f"{name}.fc2.weight"])
    return torch.sigmoid(mlp(F.adaptive_avg_pool2d(x, 1)) + mlp(F.adaptive_max_pool2d(x, 1)))


def snunet_forward(sd: SD, xA: torch.Tensor, xB: torch.Tensor) -> torch.Tensor:
    """SNUNet_ECAM.forward, models/SNUNet.py:116-152."""
    pool = lambda t: F.max_pool2d(t, 2, 2)  # noqa: E731
    x0_0A = _nested(sd, "conv0_0", xA)
    x1_0A = _nested(sd, "conv1_0", pool(x0_0A))
    x2_0A = _nested(sd, "conv2_0", pool(x1_0A))
    x3_0A = _nested(sd, "conv3_0", pool(x2_0A))
    x0_0B = _nested(sd, "conv0_0", xB)
    x1_0B = _nested(sd, "conv1_0", pool(x0_0B))
    x2_0B = _nested(sd, "conv2_0", pool(x1_0B))
    x3_0B = _nested(sd, "conv3_0", pool(x2_0B))
    x4_0B = _nested(sd, "conv4_0", pool(x3_0B))     # stream A's level 4 is commented out upstream (:123)

    x0_1 = _nested(sd, "conv0_1", torch.cat([x0_0A, x0_0B, _up(sd, "Up1_0", x1_0B)], 1))
    x1_1 = _nested(sd, "conv1_1", torch.cat([x1_0A, x1_0B, _up(sd, "Up2_0", x2_0B)], 1))
    x0_2 = _nested(sd, "conv0_2", torch.cat([x0_0A, x0_0B, x0_1, _up(sd, "Up1_1", x1_1)], 1))
    x2_1 = _nested(sd, "conv2_1", torch.cat([x2_0A, x2_0B, _up(sd, "Up3_0", x3_0B)], 1))
    x1_2 = _nested(sd, "conv1_2", torch.cat([x1_0A, x1_0B, x1_1, _up(sd, "Up2_1", x2_1)], 1))
    x0_3 = _nested(sd, "conv0_3", torch.cat([x0_0A, x0_0B, x0_1, x0_2, _up(sd, "Up1_2", x1_2)], 1))
    x3_1 = _nested(sd, "conv3_1", torch.cat([x3_0A, x3_0B, _up(sd, "Up4_0", x4_0B)], 1))
    x2_2 = _nested(sd, "conv2_2", torch.cat([x2_0A, x2_0B, x2_1, _up(sd, "Up3_1", x3_1)], 1))
    x1_3 = _nested(sd, "conv1_3", torch.cat([x1_0A, x1_0B, x1_1, x1_2, _up(sd, "Up2_2", x2_2)], 1))
    x0_4 = _nested(sd, "conv0_4", torch.cat([x0_0A, x0_0B, x0_1, x0_2, x0_3, _up(sd, "Up1_3", x1_3)], 1))

    out = torch.cat([x0_1, x0_2, x0_3, x0_4], 1)
    intra = x0_1 + x0_2 + x0_3 + x0_4
    ca1 = _channel_attention(sd, "ca1", intra)
    out = _channel_attention(sd, "ca", out) * (out + ca1.repeat(1, 4, 1, 1))
    return F.conv2d(out, sd["conv_final.weight"], sd["conv_final.bias"])


# ------------------------------------------------------------------------------------------
# smp SegCD (Siamese Unet, ResNet BasicBlock encoder)
def _basic_block(sd: SD, pre: str, x: torch.Tensor, stride: int) -> torch.Tensor:
    """torchvision BasicBlock.forward (== models/resnet.py:57-75)."""
    out = F.relu(_bn(sd, f"{pre}.bn1", F.conv2d(x, sd[f"{pre}.conv1.weight"], None, stride=stride, padding=1)))
    out = _bn(sd, f"{pre}.bn2", F.conv2d(out, sd[f"{pre}.conv2.weight"], None, padding=1))
    if f"{pre}.downsample.0.weight" in sd:
        x = _bn(sd, f"{pre}.downsample.1", F.conv2d(x, sd[f"{pre}.downsample.0.weight"], None, stride=stride))
    return F.relu(out + x)


def _bottleneck(sd: SD, pre: str, x: torch.Tensor, stride: int) -> torch.Tensor:
    """torchvision Bottleneck.forward (== models/resnet.py:104-124): the stride sits on the 3x3 conv2."""
    out = F.relu(_bn(sd, f"{pre}.bn1", F.conv2d(x, sd[f"{pre}.conv1.weight"], None)))
    out = F.relu(_bn(sd, f"{pre}.bn2", F.conv2d(out, sd[f"{pre}.conv2.weight"], None, stride=stride, padding=1)))
    out = _bn(sd, f"{pre}.bn3", F.conv2d(out, sd[f"{pre}.conv3.weight"], None))
    if f"{pre}.downsample.0.weight" in sd:
        x = _bn(sd, f"{pre}.downsample.1", F.conv2d(x, sd[f"{pre}.downsample.0.weight"], None, stride=stride))
    return F.relu(out + x)


def _resnet_features(sd: SD, x: torch.Tensor, layers) -> List[torch.Tensor]:
    """ResNetEncoder.forward, segmentation_models_pytorch/encoders/resnet.py:47-65."""
    feats = [x]
    x = F.relu(_bn(sd, "encoder.bn1", F.conv2d(x, sd["encoder.conv1.weight"], None, stride=2, padding=3)))
    feats.append(x)
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    for li, n in enumerate(layers):
        for b in range(n):
            pre = f"encoder.layer{li + 1}.{b}"
            block = _bottleneck if f"{pre}.conv3.weight" in sd else _basic_block
            x = block(sd, pre, x, 2 if (b == 0 and li > 0) else 1)
        feats.append(x)
    return feats


def _unet_decoder(sd: SD, feats: List[torch.Tensor], n_blocks: int) -> torch.Tensor:
    """UnetDecoder.forward + DecoderBlock.forward, decoders/unet/decoder.py:108-123,35-43."""
    feats = feats[1:][::-1]
    x, skips = feats[0], feats[1:]
    for i in range(n_blocks):
        x = F.interpolate(x, scale_factor=2, mode="nearest")
        if i < len(skips):
            x = torch.cat([x, skips[i]], dim=1)
        for c in ("conv1", "conv2"):
            pre = f"decoder.blocks.{i}.{c}"
            x = F.relu(_bn(sd, f"{pre}.1", F.conv2d(x, sd[f"{pre}.0.weight"], None, padding=1)))
    return x


def ffctlcd_forward(sd: SD, A: torch.Tensor, B: torch.Tensor, layers=(3, 4, 6, 3)):
    """FFCTLCD.forward, segmentation_models_pytorch/decoders/unet/model.py:407-423."""
    n_blocks = sum(1 for k in sd if k.startswith("decoder.blocks.") and k.endswith(".conv1.0.weight"))
    f1, f2 = _resnet_features(sd, A, layers), _resnet_features(sd, B, layers)
    head = lambda t: F.conv2d(t, sd["segmentation_head.0.weight"], sd["segmentation_head.0.bias"], padding=1)  # noqa: E731
    diffea = head(_unet_decoder(sd, [torch.abs(a - b) for a, b in zip(f1, f2)], n_blocks))
    m1, m2 = head(_unet_decoder(sd, f1, n_blocks)), head(_unet_decoder(sd, f2, n_blocks))
    return m1, m2, torch.min(diffea, torch.abs(m1 - m2))


def segcd_forward(sd: SD, A: torch.Tensor, B: torch.Tensor, layers=(3, 4, 6, 3)):
    """SegCD.forward, segmentation_models_pytorch/decoders/unet/model.py:316-332 -> (mask_t1, mask_t2, change)."""
    n_blocks = sum(1 for k in sd if k.startswith("decoder.blocks.") and k.endswith(".conv1.0.weight"))
    d1 = _unet_decoder(sd, _resnet_features(sd, A, layers), n_blocks)
    d2 = _unet_decoder(sd, _resnet_features(sd, B, layers), n_blocks)
    head = lambda t: F.conv2d(t, sd["segmentation_head.0.weight"], sd["segmentation_head.0.bias"], padding=1)  # noqa: E731
    m1, m2 = head(d1), head(d2)
    change = torch.min(head(torch.abs(d1 - d2)), torch.abs(m1 - m2))
    return m1, m2, change


# ------------------------------------------------------------------------------------------
# ChangeGNNV1 (ViG pyramid encoder + multi-scale difference decoder)
def _conv_bn(sd: SD, pre: str, x: torch.Tensor, stride: int = 1, padding: int = 0) -> torch.Tensor:
    """nn.Sequential(Conv2d, BatchNorm2d) as registered under `pre`.0 / `pre`.1."""
    return _bn(sd, f"{pre}.1", F.conv2d(x, sd[f"{pre}.0.weight"], sd.get(f"{pre}.0.bias"), stride=stride, padding=padding))


def _grapher(sd: SD, pre: str, x: torch.Tensor, k: int, dilation: int, r: int) -> torch.Tensor:
    """gcn_lib Grapher.forward (absent from the reference tree: SURVEY App. D, oracle/gcn_lib_restated.py)."""
    from oracle import gcn
    t = x
    x = _conv_bn(sd, f"{pre}.fc1", x)
    b, c, h, w = x.shape
    y = F.avg_pool2d(x, r, r).reshape(b, c, -1, 1) if r > 1 else None
    xn = x.reshape(b, c, -1, 1)
    e = gcn.dense_dilated_knn_graph(xn, y, k, dilation, sd.get(f"{pre}.relative_pos"))
    z = gcn.mr_features(xn, e, y)
    g = f"{pre}.graph_conv.gconv.nn"
    z = F.gelu(_bn(sd, f"{g}.1", F.conv2d(z, sd[f"{g}.0.weight"], sd.get(f"{g}.0.bias"), groups=4)))
    x = _conv_bn(sd, f"{pre}.fc2", z.reshape(b, 2 * c, h, w))
    return x + t


def _ffn(sd: SD, pre: str, x: torch.Tensor) -> torch.Tensor:
    """models/pyramid_vig.py:41-63."""
    return _conv_bn(sd, f"{pre}.fc2", F.gelu(_conv_bn(sd, f"{pre}.fc1", x))) + x


def vig_encoder_features(sd: SD, x: torch.Tensor, blocks=(2, 2, 6, 2), k: int = 9, pre: str = "encoder") -> List[torch.Tensor]:
    """EncoderV1.forward_features, models/ChangeVIG.py:85-94 (Stem / Downsample: pyramid_vig.py:66-100)."""
    s = f"{pre}.stem.convs"
    x = F.gelu(_bn(sd, f"{s}.1", F.conv2d(x, sd[f"{s}.0.weight"], sd[f"{s}.0.bias"], stride=2, padding=1)))
    x = F.gelu(_bn(sd, f"{s}.4", F.conv2d(x, sd[f"{s}.3.weight"], sd[f"{s}.3.bias"], stride=2, padding=1)))
    x = _bn(sd, f"{s}.7", F.conv2d(x, sd[f"{s}.6.weight"], sd[f"{s}.6.bias"], padding=1))
    x = x + sd[f"{pre}.pos_embed"]
    max_dilation = 49 // k
    reduce_ratios = (4, 2, 1, 1)
    outs, bi, idx = [], 0, 0
    for i, n in enumerate(blocks):
        if i > 0:
            x = _conv_bn(sd, f"{pre}.backbone.{bi}.conv", x, stride=2, padding=1)
            bi += 1
        for _ in range(n):
            x = _grapher(sd, f"{pre}.backbone.{bi}.0", x, k, min(idx // 4 + 1, max_dilation), reduce_ratios[i])
            x = _ffn(sd, f"{pre}.backbone.{bi}.1", x)
            bi += 1
            idx += 1
        outs.append(x)
    return outs


def changegnn_decoder(sd: SD, f1: List[torch.Tensor], f2: List[torch.Tensor], pre: str = "decoder") -> List[torch.Tensor]:
    """DecoderV1.forward with decoder_heads="MLP", models/ChangeVIG.py:192-281."""
    def mlp(k, t):          # MLP: per-pixel Linear (ChangeFormer.py:677-688)
        return F.conv2d(t, sd[f"{pre}.decoder_heads_c{k}.proj.weight"][:, :, None, None], sd[f"{pre}.decoder_heads_c{k}.proj.bias"])

    def diff(k, t):         # conv_diff: conv, PReLU, BN, (dropout,) conv, PReLU, BN (ChangeFormer.py:1138-1148)
        d = f"{pre}.diff_c{k}"
        t = _bn(sd, f"{d}.2", F.prelu(F.conv2d(t, sd[f"{d}.0.weight"], sd[f"{d}.0.bias"], padding=1), sd[f"{d}.1.weight"]))
        return _bn(sd, f"{d}.6", F.prelu(F.conv2d(t, sd[f"{d}.4.weight"], sd[f"{d}.4.bias"], padding=1), sd[f"{d}.5.weight"]))

    def pred(k, t):         # make_prediction: conv, ReLU, BN, conv (ChangeFormer.py:1151-1157)
        m = f"{pre}.make_pred_c{k}"
        t = _bn(sd, f"{m}.2", F.relu(F.conv2d(t, sd[f"{m}.0.weight"], sd[f"{m}.0.bias"], padding=1)))
        return F.conv2d(t, sd[f"{m}.3.weight"], sd[f"{m}.3.bias"], padding=1)

    def resblock(name, t):  # ResidualBlock (ChangeFormerBaseNetworks.py:108-120)
        o = F.relu(F.conv2d(t, sd[f"{name}.conv1.conv2d.weight"], sd[f"{name}.conv1.conv2d.bias"], padding=1))
        return F.conv2d(o, sd[f"{name}.conv2.conv2d.weight"], sd[f"{name}.conv2.conv2d.bias"], padding=1) * 0.1 + t

    size = f1[0].shape[2:]
    outs, ups, c_prev = [], [], None
    for k in (4, 3, 2, 1):
        a, b = f1[k - 1], f2[k - 1]
        c = diff(k, torch.cat((mlp(k, a), mlp(k, b)), dim=1))
        if c_prev is not None:
            c = c + F.interpolate(c_prev, scale_factor=2, mode="bilinear")
        outs.append(pred(k, c))
        ups.append(c if k == 1 else F.interpolate(c, size=size, mode="bilinear", align_corners=False))
        c_prev = c
    x = _conv_bn(sd, f"{pre}.linear_fuse", torch.cat(ups, dim=1))
    x = F.conv_transpose2d(x, sd[f"{pre}.convd2x.conv2d.weight"], sd[f"{pre}.convd2x.conv2d.bias"], stride=2, padding=1)
    x = resblock(f"{pre}.dense_2x.0", x)
    x = F.conv_transpose2d(x, sd[f"{pre}.convd1x.conv2d.weight"], sd[f"{pre}.convd1x.conv2d.bias"], stride=2, padding=1)
    x = resblock(f"{pre}.dense_1x.0", x)
    outs.append(F.conv2d(x, sd[f"{pre}.change_probability.conv2d.weight"], sd[f"{pre}.change_probability.conv2d.bias"], padding=1))
    return outs


def changegnn_forward(sd: SD, x1: torch.Tensor, x2: torch.Tensor) -> List[torch.Tensor]:
    """ChangeGNNV1.forward, models/ChangeVIG.py:309-312: list of 5 tensors, full-resolution logits last."""
    return changegnn_decoder(sd, vig_encoder_features(sd, x1), vig_encoder_features(sd, x2))


# ------------------------------------------------------------------------------------------
# ChangeFormerV6 (MiT-style Siamese transformer encoder + the same multi-scale difference decoder)
def _ln(sd: SD, name: str, x: torch.Tensor, eps: float) -> torch.Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[f"{name}.weight"], sd[f"{name}.bias"], eps)


def _mit_attention(sd: SD, pre: str, x: torch.Tensor, h: int, w: int, heads: int, sr: int) -> torch.Tensor:
    """Attention.forward, models/ChangeFormer.py:338-358 (attn_drop / proj_drop are identities in eval)."""
    b, n, c = x.shape
    d = c // heads
    q = F.linear(x, sd[f"{pre}.q.weight"], sd[f"{pre}.q.bias"]).reshape(b, n, heads, d).permute(0, 2, 1, 3)
    if sr > 1:
        x_ = x.permute(0, 2, 1).reshape(b, c, h, w)
        x_ = F.conv2d(x_, sd[f"{pre}.sr.weight"], sd[f"{pre}.sr.bias"], stride=sr).reshape(b, c, -1).permute(0, 2, 1)
        x_ = _ln(sd, f"{pre}.norm", x_, 1e-5)              # nn.LayerNorm(dim): default eps (:316)
    else:
        x_ = x
    kv = F.linear(x_, sd[f"{pre}.kv.weight"], sd[f"{pre}.kv.bias"]).reshape(b, -1, 2, heads, d).permute(2, 0, 3, 1, 4)
    k, v = kv[0], kv[1]
    attn = ((q @ k.transpose(-2, -1)) * (d ** -0.5)).softmax(dim=-1)
    x = (attn @ v).transpose(1, 2).reshape(b, n, c)
    return F.linear(x, sd[f"{pre}.proj.weight"], sd[f"{pre}.proj.bias"])


def _mit_mlp(sd: SD, pre: str, x: torch.Tensor, h: int, w: int) -> torch.Tensor:
    """Mlp.forward + DWConv, models/ChangeFormer.py:283-291,512-523."""
    b, n, _ = x.shape
    x = F.linear(x, sd[f"{pre}.fc1.weight"], sd[f"{pre}.fc1.bias"])
    c = x.shape[2]
    x = x.transpose(1, 2).reshape(b, c, h, w)
    x = F.conv2d(x, sd[f"{pre}.dwconv.dwconv.weight"], sd[f"{pre}.dwconv.dwconv.bias"], padding=1, groups=c)
    x = F.gelu(x.flatten(2).transpose(1, 2))
    return F.linear(x, sd[f"{pre}.fc2.weight"], sd[f"{pre}.fc2.bias"])


def mit_encoder_features(sd: SD, x: torch.Tensor, pre: str = "Tenc_x2", depths=(3, 3, 4, 3), heads=(1, 2, 4, 8),
                         srs=(8, 4, 2, 1)) -> List[torch.Tensor]:
    """EncoderTransformer_v3.forward_features, models/ChangeFormer.py:1434-1469 (ChangeFormerV6: patch 7 at every stage,
    block LayerNorm eps 1e-6, :1681-1683; OverlapPatchEmbed's own LayerNorm keeps the default eps, :226)."""
    outs = []
    for s in range(4):
        pe = f"{pre}.patch_embed{s + 1}"
        x = F.conv2d(x, sd[f"{pe}.proj.weight"], sd[f"{pe}.proj.bias"], stride=4 if s == 0 else 2, padding=sd[f"{pe}.proj.weight"].shape[2] // 2)
        b, c, h, w = x.shape
        t = _ln(sd, f"{pe}.norm", x.flatten(2).transpose(1, 2), 1e-5)
        for i in range(depths[s]):
            blk = f"{pre}.block{s + 1}.{i}"
            t = t + _mit_attention(sd, f"{blk}.attn", _ln(sd, f"{blk}.norm1", t, 1e-6), h, w, heads[s], srs[s])
            t = t + _mit_mlp(sd, f"{blk}.mlp", _ln(sd, f"{blk}.norm2", t, 1e-6), h, w)
        t = _ln(sd, f"{pre}.norm{s + 1}", t, 1e-6)
        x = t.reshape(b, h, w, -1).permute(0, 3, 1, 2).contiguous()
        outs.append(x)
    return outs


def changeformer_decoder(sd: SD, f1: List[torch.Tensor], f2: List[torch.Tensor], pre: str = "TDec_x2") -> List[torch.Tensor]:
    """DecoderTransformer_v3.forward, models/ChangeFormer.py:1558-1631: DecoderV1's wiring with heads named linear_c{k}."""
    renamed = {k.replace(f"{pre}.linear_c", f"{pre}.decoder_heads_c"): v for k, v in sd.items()}
    return changegnn_decoder(renamed, f1, f2, pre)


def changeformer_forward(sd: SD, x1: torch.Tensor, x2: torch.Tensor) -> List[torch.Tensor]:
    """ChangeFormerV6.forward, models/ChangeFormer.py:1691-1701: list of 5 tensors, full-resolution logits last."""
    return changeformer_decoder(sd, mit_encoder_features(sd, x1), mit_encoder_features(sd, x2))


# ------------------------------------------------------------------------------------------
# DTCDSCN (CDNet34: SE-ResNet-34 Siamese encoder, dilated centre block, SCSE decoder on the feature differences)
def _se_block(sd: SD, pre: str, x: torch.Tensor, stride: int) -> torch.Tensor:
    """SEBasicBlock.forward, models/DTCDSCN.py:93-109 (SELayer :22-26)."""
    out = F.relu(_bn(sd, f"{pre}.bn1", F.conv2d(x, sd[f"{pre}.conv1.weight"], None, stride=stride, padding=1)))
    out = _bn(sd, f"{pre}.bn2", F.conv2d(out, sd[f"{pre}.conv2.weight"], None, padding=1))
    y = out.mean(dim=(2, 3))
    y = torch.sigmoid(F.linear(F.relu(F.linear(y, sd[f"{pre}.se.fc.0.weight"])), sd[f"{pre}.se.fc.2.weight"]))
    out = out * y[:, :, None, None]
    if f"{pre}.downsample.0.weight" in sd:
        x = _bn(sd, f"{pre}.downsample.1", F.conv2d(x, sd[f"{pre}.downsample.0.weight"], None, stride=stride))
    return F.relu(out + x)


def _dtcdscn_encoder(sd: SD, x: torch.Tensor, layers=(3, 4, 6, 3)) -> List[torch.Tensor]:
    """CDNet_model.forward, encoder half (models/DTCDSCN.py:246-254)."""
    x = F.relu(_bn(sd, "firstbn", F.conv2d(x, sd["firstconv.weight"], None, stride=2, padding=3)))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    feats = []
    for li, n in enumerate(layers):
        for b in range(n):
            x = _se_block(sd, f"encoder{li + 1}.{b}", x, 2 if (b == 0 and li > 0) else 1)
        feats.append(x)
    return feats


def _scse(sd: SD, pre: str, x: torch.Tensor) -> torch.Tensor:
    """SCSEBlock.forward, models/DTCDSCN.py:164-173."""
    c = F.adaptive_avg_pool2d(x, 1)
    c = torch.sigmoid(F.conv2d(F.relu(F.conv2d(c, sd[f"{pre}.channel_excitation.0.weight"])), sd[f"{pre}.channel_excitation.2.weight"]))
    s = torch.sigmoid(F.conv2d(x, sd[f"{pre}.spatial_se.0.weight"]))
    return x * c + x * s


def _dtcdscn_decoder_block(sd: SD, pre: str, x: torch.Tensor) -> torch.Tensor:
    """DecoderBlock.forward, models/DTCDSCN.py:129-141."""
    x = F.relu(_bn(sd, f"{pre}.norm1", F.conv2d(x, sd[f"{pre}.conv1.weight"], sd[f"{pre}.conv1.bias"])))
    x = x + _scse(sd, f"{pre}.scse", x)
    x = F.relu(_bn(sd, f"{pre}.norm2", F.conv_transpose2d(x, sd[f"{pre}.deconv2.weight"], sd[f"{pre}.deconv2.bias"], stride=2, padding=1,
                                                          output_padding=1)))
    return F.relu(_bn(sd, f"{pre}.norm3", F.conv2d(x, sd[f"{pre}.conv3.weight"], sd[f"{pre}.conv3.bias"])))


def dtcdscn_forward(sd: SD, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """CDNet_model.forward, models/DTCDSCN.py:244-313 (the change branch: `*_master` modules on feature differences)."""
    ex, ey = _dtcdscn_encoder(sd, x), _dtcdscn_encoder(sd, y)
    t = ex[3] - ey[3]
    d = t                                              # Dblock.forward (:65-71): x + sum of the dilated chain
    acc = t
    for i, dil in enumerate((1, 2, 4, 8)):
        d = F.relu(F.conv2d(d, sd[f"dblock_master.dilate{i + 1}.weight"], sd[f"dblock_master.dilate{i + 1}.bias"], padding=dil, dilation=dil))
        acc = acc + d
    d4 = _dtcdscn_decoder_block(sd, "decoder4_master", acc) + ex[2] - ey[2]
    d3 = _dtcdscn_decoder_block(sd, "decoder3_master", d4) + ex[1] - ey[1]
    d2 = _dtcdscn_decoder_block(sd, "decoder2_master", d3) + ex[0] - ey[0]
    d1 = _dtcdscn_decoder_block(sd, "decoder1_master", d2)
    out = F.relu(F.conv_transpose2d(d1, sd["finaldeconv1_master.weight"], sd["finaldeconv1_master.bias"], stride=2, padding=1))
    out = F.relu(F.conv2d(out, sd["finalconv2_master.weight"], sd["finalconv2_master.bias"], padding=1))
    return F.conv2d(out, sd["finalconv3_master.weight"], sd["finalconv3_master.bias"], padding=1)


# ------------------------------------------------------------------------------------------
# BIT (models/networks.py: ResNet :223-305, BASE_Transformer :308-441; blocks in models/help_funcs.py)
def _bit_backbone(sd: SD, x: torch.Tensor, stages: int) -> torch.Tensor:
    """ResNet.forward_single, models/networks.py:277-305.  The backbone is models/resnet.py's resnet18 with
    replace_stride_with_dilation=[False, True, True]: layer3 / layer4 keep stride 1, and BasicBlock silently resets the
    requested dilation to 1 (models/resnet.py:47-49), so they are plain stride-1 3x3 blocks."""
    x = F.relu(_bn(sd, "resnet.bn1", F.conv2d(x, sd["resnet.conv1.weight"], None, stride=2, padding=3)))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    for li in range(min(stages, 5) - 1):
        for b in range(2):
            x = _basic_block(sd, f"resnet.layer{li + 1}.{b}", x, 2 if (li == 1 and b == 0) else 1)
    x = F.interpolate(x, scale_factor=2, mode="nearest")               # self.upsamplex2 (if_upsample_2x=True)
    return F.conv2d(x, sd["conv_pred.weight"], sd["conv_pred.bias"], padding=1)


def _bit_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int, scale: float, softmax: bool = True) -> torch.Tensor:
    """help_funcs.py Attention / Cross_Attention core (:87-110, :122-146): q [b, n, inner], k / v [b, m, inner]."""
    b, n, inner = q.shape
    d = inner // heads
    qh, kh, vh = (t.reshape(b, -1, heads, d).transpose(1, 2) for t in (q, k, v))
    dots = torch.einsum("bhid,bhjd->bhij", qh, kh) * scale
    attn = dots.softmax(dim=-1) if softmax else dots
    return torch.einsum("bhij,bhjd->bhid", attn, vh).transpose(1, 2).reshape(b, n, inner)


def _bit_ff(sd: SD, pre: str, x: torch.Tensor) -> torch.Tensor:
    """Residual(PreNorm(FeedForward)), help_funcs.py:21-26,37-43,56-68."""
    y = F.layer_norm(x, (x.shape[-1],), sd[f"{pre}.fn.norm.weight"], sd[f"{pre}.fn.norm.bias"])
    y = F.gelu(F.linear(y, sd[f"{pre}.fn.fn.net.0.weight"], sd[f"{pre}.fn.fn.net.0.bias"]))
    return F.linear(y, sd[f"{pre}.fn.fn.net.3.weight"], sd[f"{pre}.fn.fn.net.3.bias"]) + x


def bit_forward(sd: SD, x1: torch.Tensor, x2: torch.Tensor, stages: int = 4, heads: int = 8, decoder_softmax: bool = True) -> torch.Tensor:
    """BASE_Transformer.forward, models/networks.py:405-441 (tokenizer=True, token_trans=True, with_pos='learned',
    with_decoder=True, with_decoder_pos=None: the registry's three BIT keys, :174-182); without the transformer
    parameters in `sd` it is ResNet.forward (:263-275, key 'base_resnet18').  Returns the logits (the reference wraps
    them in a one-element list for BASE_Transformer)."""
    f1, f2 = _bit_backbone(sd, x1, stages), _bit_backbone(sd, x2, stages)
    if "conv_a.weight" in sd:
        b, c, h, w = f1.shape
        dim = c

        def tokens(f):                                               # _forward_semantic_tokens, :359-367
            a = F.conv2d(f, sd["conv_a.weight"]).reshape(b, -1, h * w).softmax(dim=-1)
            return torch.einsum("bln,bcn->blc", a, f.reshape(b, c, h * w))

        t = torch.cat([tokens(f1), tokens(f2)], dim=1) + sd["pos_embedding"]          # :419-420, 380-384
        n_enc = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("transformer.layers."))
        for l in range(n_enc):                                       # Transformer.forward, help_funcs.py:159-163
            pre = f"transformer.layers.{l}"
            y = F.layer_norm(t, (dim,), sd[f"{pre}.0.fn.norm.weight"], sd[f"{pre}.0.fn.norm.bias"])
            q, k, v = F.linear(y, sd[f"{pre}.0.fn.fn.to_qkv.weight"]).chunk(3, dim=-1)
            o = _bit_attention(q, k, v, heads, dim ** -0.5)
            t = F.linear(o, sd[f"{pre}.0.fn.fn.to_out.0.weight"], sd[f"{pre}.0.fn.fn.to_out.0.bias"]) + t
            t = _bit_ff(sd, f"{pre}.1", t)
        t1, t2 = t.chunk(2, dim=1)
        n_dec = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("transformer_decoder.layers."))

        def decode(f, m):                                            # _forward_transformer_decoder, :386-394
            x = f.reshape(b, c, h * w).transpose(1, 2)
            for l in range(n_dec):                                   # TransformerDecoder.forward, help_funcs.py:177-182
                pre = f"transformer_decoder.layers.{l}"
                nw, nb = sd[f"{pre}.0.fn.norm.weight"], sd[f"{pre}.0.fn.norm.bias"]
                xn, mn = F.layer_norm(x, (dim,), nw, nb), F.layer_norm(m, (dim,), nw, nb)        # PreNorm2: one norm for both
                q = F.linear(xn, sd[f"{pre}.0.fn.fn.to_q.weight"])
                k = F.linear(mn, sd[f"{pre}.0.fn.fn.to_k.weight"])
                v = F.linear(mn, sd[f"{pre}.0.fn.fn.to_v.weight"])
                o = _bit_attention(q, k, v, heads, dim ** -0.5, decoder_softmax)
                x = F.linear(o, sd[f"{pre}.0.fn.fn.to_out.0.weight"], sd[f"{pre}.0.fn.fn.to_out.0.bias"]) + x
                x = _bit_ff(sd, f"{pre}.1", x)
            return x.transpose(1, 2).reshape(b, c, h, w)

        f1, f2 = decode(f1, t1), decode(f2, t2)
    x = torch.abs(f1 - f2)
    x = F.interpolate(x, scale_factor=4, mode="bilinear")           # self.upsamplex4 (align_corners unset -> False)
    x = F.relu(_bn(sd, "classifier.1", F.conv2d(x, sd["classifier.0.weight"], None, padding=1)))
    return F.conv2d(x, sd["classifier.3.weight"], sd["classifier.3.bias"], padding=1)


# ------------------------------------------------------------------------------------------
# IFNet / DSIFN (models/DSIFN.py): shared VGG16 features, deeply supervised difference decoder with channel / spatial attention
_VGG_CONVS = (0, 2, 5, 7, 10, 12, 14, 17, 19, 21, 24, 26, 28)        # torchvision vgg16().features[:30]: conv indices
_VGG_POOLS = (4, 9, 16, 23)
_VGG_TAPS = (3, 8, 15, 22, 29)


def _vgg16_features(sd: SD, pre: str, x: torch.Tensor) -> List[torch.Tensor]:
    """vgg16_base.forward, models/DSIFN.py:15-21: outputs of features[3, 8, 15, 22, 29] (the ReLUs closing each block)."""
    outs = []
    for i in range(30):
        if i in _VGG_CONVS:
            x = F.relu(F.conv2d(x, sd[f"{pre}.features.{i}.weight"], sd[f"{pre}.features.{i}.bias"], padding=1))
        elif i in _VGG_POOLS:
            x = F.max_pool2d(x, kernel_size=2, stride=2)
        if i in _VGG_TAPS:
            outs.append(x)
    return outs


def _dsifn_conv_bn(sd: SD, pre: str, x: torch.Tensor) -> torch.Tensor:
    """conv2d_bn, models/DSIFN.py:54-60: conv3x3 -> PReLU -> BatchNorm (-> Dropout: identity in eval mode)."""
    x = F.prelu(F.conv2d(x, sd[f"{pre}.0.weight"], sd[f"{pre}.0.bias"], padding=1), sd[f"{pre}.1.weight"])
    return _bn(sd, f"{pre}.2", x)


def _dsifn_ca(sd: SD, pre: str, x: torch.Tensor) -> torch.Tensor:
    """ChannelAttention.forward, models/DSIFN.py:32-36."""
    def fc(v):
        return F.conv2d(F.relu(F.conv2d(v, sd[f"{pre}.fc1.weight"])), sd[f"{pre}.fc2.weight"])
    return torch.sigmoid(fc(F.adaptive_avg_pool2d(x, 1)) + fc(F.adaptive_max_pool2d(x, 1)))


def _dsifn_sa(sd: SD, pre: str, x: torch.Tensor) -> torch.Tensor:
    """SpatialAttention.forward, models/DSIFN.py:45-51."""
    m = torch.cat([x.mean(dim=1, keepdim=True), x.max(dim=1, keepdim=True)[0]], dim=1)
    return torch.sigmoid(F.conv2d(m, sd[f"{pre}.conv1.weight"], None, padding=3))


def dsifn_forward(sd: SD, t1: torch.Tensor, t2: torch.Tensor) -> torch.Tensor:
    """DSIFN.forward, models/DSIFN.py:119-188: returns `out` = o5_conv4(...) (the four deep-supervision sigmoids are appended
    to a list the reference discards)."""
    f1, f2 = _vgg16_features(sd, "t1_base", t1), _vgg16_features(sd, "t2_base", t2)
    x = torch.cat((f1[4], f2[4]), dim=1)
    x = _dsifn_conv_bn(sd, "o1_conv2", _dsifn_conv_bn(sd, "o1_conv1", x))
    x = _bn(sd, "bn_sa1", _dsifn_sa(sd, "sa1", x) * x)
    for b, n_convs in ((2, 3), (3, 3), (4, 3), (5, 3)):
        x = F.conv_transpose2d(x, sd[f"trans_conv{b - 1}.weight"], sd[f"trans_conv{b - 1}.bias"], stride=2)
        x = torch.cat((x, f1[5 - b], f2[5 - b]), dim=1)
        x = _dsifn_ca(sd, f"ca{b}", x) * x
        for i in range(n_convs):
            x = _dsifn_conv_bn(sd, f"o{b}_conv{i + 1}", x)
        x = _bn(sd, f"bn_sa{b}", _dsifn_sa(sd, f"sa{b}", x) * x)
    return F.conv2d(x, sd["o5_conv4.weight"], sd["o5_conv4.bias"])


# ------------------------------------------------------------------------------------------
# ChangeGNNV2 / ChangeGNNV2_Compare (models/ChangeVIG.py:315-460,537-918): EncoderV2 == EncoderV1; DecoderV2 = HFFM + VFFM
def _seq_conv_bn(sd: SD, pre: str, i: int, x: torch.Tensor, padding: int = 0) -> torch.Tensor:
    """nn.Sequential members i (Conv2d) and i + 1 (BatchNorm2d)."""
    return _bn(sd, f"{pre}.{i + 1}", F.conv2d(x, sd[f"{pre}.{i}.weight"], sd.get(f"{pre}.{i}.bias"), padding=padding))


def _res_bottleneck(sd: SD, pre: str, out: torch.Tensor) -> torch.Tensor:
    """act(conv_res(out) + conv(out)) shared by Cross_ConCat / Sub / Abs / Conc (ChangeVIG.py:323-347,670-750)."""
    r = _seq_conv_bn(sd, f"{pre}.conv_res", 0, out, padding=1)
    c = F.relu(_seq_conv_bn(sd, f"{pre}.conv", 0, out))
    c = F.relu(_seq_conv_bn(sd, f"{pre}.conv", 3, c, padding=1))
    c = _seq_conv_bn(sd, f"{pre}.conv", 6, c)
    return F.relu(r + c)


def _cross_concat_v2(sd: SD, pre: str, a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Cross_ConCat.forward, ChangeVIG.py:339-347: channels interleaved (a0, b0, a1, b1, ...), grouped 3x3 conv (one group per
    channel pair), BN, ReLU, then the residual bottleneck."""
    n, c, h, w = a.shape
    z = torch.stack([a, b], dim=2).reshape(n, 2 * c, h, w)
    out = F.relu(_bn(sd, f"{pre}.diff.1", F.conv2d(z, sd[f"{pre}.diff.0.weight"], sd[f"{pre}.diff.0.bias"], padding=1, groups=c)))
    return _res_bottleneck(sd, pre, out)


def _global_local(sd: SD, pre: str, x: torch.Tensor) -> torch.Tensor:
    """Global_Local.forward, ChangeVIG.py:377-391."""
    c = x.shape[1]
    pooled = torch.cat([F.adaptive_avg_pool2d(x, 1), F.adaptive_max_pool2d(x, 1)], dim=2)            # [n, c, 2, 1]
    ch = F.relu(_bn(sd, f"{pre}.channel_bn", F.conv2d(pooled, sd[f"{pre}.channel_conv.weight"], sd[f"{pre}.channel_conv.bias"], groups=c)))
    sp_in = torch.cat([x.mean(dim=1, keepdim=True), x.max(dim=1, keepdim=True)[0]], dim=1)
    sp = F.relu(F.conv2d(sp_in, sd[f"{pre}.spatial_conv.weight"], sd[f"{pre}.spatial_conv.bias"], padding=2))
    gated = torch.sigmoid(ch * sp) * x
    loc = torch.cat([F.conv2d(x, sd[f"{pre}.local_conv1.weight"], sd[f"{pre}.local_conv1.bias"], groups=c),
                     F.conv2d(x, sd[f"{pre}.local_conv2.weight"], sd[f"{pre}.local_conv2.bias"], padding=1, groups=c),
                     F.conv2d(x, sd[f"{pre}.local_conv3.weight"], sd[f"{pre}.local_conv3.bias"], padding=3, groups=c)], dim=1)
    loc = F.conv2d(loc, sd[f"{pre}.local_conv4.weight"], sd[f"{pre}.local_conv4.bias"])
    loc = F.conv2d(F.relu(_bn(sd, f"{pre}.local_bn", loc)), sd[f"{pre}.local_conv5.weight"], sd[f"{pre}.local_conv5.bias"], padding=1)
    return gated + loc


def _vffm(sd: SD, pre: str, low: torch.Tensor, high: torch.Tensor) -> torch.Tensor:
    """VFFM.forward, ChangeVIG.py:452-460."""
    high = F.conv_transpose2d(high, sd[f"{pre}.up.up.weight"], sd[f"{pre}.up.up.bias"], stride=2)
    mixed = low + high

    def mlp(p, v):                                   # Sequential: [pool,] conv, BN, ReLU, conv, BN
        o = 1 if p != "local_att" else 0
        v = F.relu(_seq_conv_bn(sd, f"{pre}.{p}", o, v))
        return _seq_conv_bn(sd, f"{pre}.{p}", o + 3, v)

    wei = torch.sigmoid(mlp("global_avg", F.adaptive_avg_pool2d(mixed, 1)) + mlp("global_max", F.adaptive_max_pool2d(mixed, 1))
                        + mlp("local_att", mixed))
    return 2 * low * wei + 2 * high * (1 - wei)


def changegnn_v2_forward(sd: SD, x1: torch.Tensor, x2: torch.Tensor, diff_mode: str = "cross") -> List[torch.Tensor]:
    """ChangeGNNV2.forward (ChangeVIG.py:661-664) / ChangeGNNV2_Compare.forward (:915-918): [full-resolution logits].
    diff_mode: 'cross' (HFFM: Cross_ConCat) or the Compare variants 'sub' / 'abs' / 'conc' (HFFM_Compare, :753-765)."""
    f1, f2 = vig_encoder_features(sd, x1), vig_encoder_features(sd, x2)
    pre = "decoder"

    def hffm(k):
        a, b = f1[k - 1], f2[k - 1]
        if diff_mode == "cross":
            d = _cross_concat_v2(sd, f"{pre}.hffm{k}.cross_conc", a, b)
        elif diff_mode == "sub":
            d = _res_bottleneck(sd, f"{pre}.hffm{k}.diff", a - b)
        elif diff_mode == "abs":
            d = _res_bottleneck(sd, f"{pre}.hffm{k}.diff", (a - b).abs())
        else:
            q = f"{pre}.hffm{k}.diff"
            out = F.relu(_bn(sd, f"{q}.diff.1", F.conv2d(torch.cat([a, b], dim=1), sd[f"{q}.diff.0.weight"], sd[f"{q}.diff.0.bias"], padding=1)))
            d = _res_bottleneck(sd, q, out)
        return _global_local(sd, f"{pre}.hffm{k}.global_local", d)

    c = _vffm(sd, f"{pre}.vffm1", hffm(1), _vffm(sd, f"{pre}.vffm2", hffm(2), _vffm(sd, f"{pre}.vffm3", hffm(3), hffm(4))))

    def resblock(q, x):                              # ResidualBlock.forward, ChangeFormerBaseNetworks.py:113-120
        r = x
        o = F.relu(F.conv2d(x, sd[f"{q}.conv1.conv2d.weight"], sd[f"{q}.conv1.conv2d.bias"], padding=1))
        return F.conv2d(o, sd[f"{q}.conv2.conv2d.weight"], sd[f"{q}.conv2.conv2d.bias"], padding=1) * 0.1 + r

    x = F.conv_transpose2d(c, sd[f"{pre}.convd2x.conv2d.weight"], sd[f"{pre}.convd2x.conv2d.bias"], stride=2, padding=1)
    x = resblock(f"{pre}.dense_2x.0", x)
    x = F.conv_transpose2d(x, sd[f"{pre}.convd1x.conv2d.weight"], sd[f"{pre}.convd1x.conv2d.bias"], stride=2, padding=1)
    x = resblock(f"{pre}.dense_1x.0", x)
    return [F.conv2d(x, sd[f"{pre}.change_probability.conv2d.weight"], sd[f"{pre}.change_probability.conv2d.bias"], padding=1)]


# ------------------------------------------------------------------------------------------
# VIG_V20_2 (registry key "GNN", models/ChangeVIG.py:921-1289): the same ViG encoder (prefix VIG_x2), conv_diff_V20 + csam_V20 + AFF
def _csam_v20(sd: SD, pre: str, x: torch.Tensor) -> torch.Tensor:
    """csam_V20.forward, ChangeVIG.py:982-994."""
    c = x.shape[1]
    pooled = torch.cat([F.adaptive_avg_pool2d(x, 1), F.adaptive_max_pool2d(x, 1)], dim=2)
    ch = F.gelu(_bn(sd, f"{pre}.batch_normal1", F.conv2d(pooled, sd[f"{pre}.conv1_1.weight"], sd[f"{pre}.conv1_1.bias"], groups=c)))
    ch = F.linear(F.relu(F.linear(ch.permute(0, 2, 3, 1), sd[f"{pre}.liner1.weight"])), sd[f"{pre}.liner2.weight"], sd[f"{pre}.liner2.bias"])
    ch = ch.permute(0, 3, 1, 2)
    sp = torch.cat([x.mean(dim=1, keepdim=True), x.max(dim=1, keepdim=True)[0]], dim=1)
    sp = F.conv2d(F.relu(F.conv2d(sp, sd[f"{pre}.conv2_1.weight"], None, padding=1)), sd[f"{pre}.conv2_2.weight"], None, padding=1)
    return _bn(sd, f"{pre}.bt", (torch.sigmoid(ch) + torch.sigmoid(sp)) * x)


def _aff(sd: SD, pre: str, x: torch.Tensor, residual: torch.Tensor) -> torch.Tensor:
    """AFF.forward, ChangeVIG.py:1019-1028."""
    xa = x + residual
    xl = _seq_conv_bn(sd, f"{pre}.local_att", 3, F.relu(_seq_conv_bn(sd, f"{pre}.local_att", 0, xa)))
    xg = _seq_conv_bn(sd, f"{pre}.global_att", 4, F.relu(_seq_conv_bn(sd, f"{pre}.global_att", 1, F.adaptive_avg_pool2d(xa, 1))))
    wei = torch.sigmoid(xl + xg)
    return 2 * x * wei + 2 * residual * (1 - wei)


def vig_v20_forward(sd: SD, x1: torch.Tensor, x2: torch.Tensor) -> List[torch.Tensor]:
    """VIG_V20_2.forward, ChangeVIG.py:1283-1289 with DecoderVIG_V20_2.forward (:1186-1239): [full-resolution logits]."""
    f1, f2 = vig_encoder_features(sd, x1, pre="VIG_x2"), vig_encoder_features(sd, x2, pre="VIG_x2")
    d = "TDec_x2"

    def scale(k):
        return _csam_v20(sd, f"{d}.csam{k}", _cross_concat_v2(sd, f"{d}.diff_c{k}", f1[k - 1], f2[k - 1]))

    def up(k, t):
        return F.conv_transpose2d(t, sd[f"{d}.trans_conv{k}.weight"], sd[f"{d}.trans_conv{k}.bias"], stride=2)

    c = up(4, scale(4))
    c = up(3, _aff(sd, f"{d}.aff3", scale(3), c))
    c = up(2, _aff(sd, f"{d}.aff2", scale(2), c))
    c = _aff(sd, f"{d}.aff1", scale(1), c)

    def resblock(q, x):
        o = F.relu(F.conv2d(x, sd[f"{q}.conv1.conv2d.weight"], sd[f"{q}.conv1.conv2d.bias"], padding=1))
        return F.conv2d(o, sd[f"{q}.conv2.conv2d.weight"], sd[f"{q}.conv2.conv2d.bias"], padding=1) * 0.1 + x

    x = F.conv_transpose2d(c, sd[f"{d}.convd2x.conv2d.weight"], sd[f"{d}.convd2x.conv2d.bias"], stride=2, padding=1)
    x = resblock(f"{d}.dense_2x.0", x)
    x = F.conv_transpose2d(x, sd[f"{d}.convd1x.conv2d.weight"], sd[f"{d}.convd1x.conv2d.bias"], stride=2, padding=1)
    x = resblock(f"{d}.dense_1x.0", x)
    return [F.conv2d(x, sd[f"{d}.change_probability.conv2d.weight"], sd[f"{d}.change_probability.conv2d.bias"], padding=1)]


# ------------------------------------------------------------------------------------------
# ChangeFormerV1 / V2 (models/ChangeFormer.py:644-674, 918-948): Tenc (EncoderTransformer, patch 7/s4 then 3/s2, depths 3-4-6-3) on both
# dates, |fx1 - fx2| per scale, then convprojection_base (V1) or TDec (V2)
def _cf_resblock(sd: SD, q: str, x: torch.Tensor) -> torch.Tensor:
    """ResidualBlock.forward, ChangeFormerBaseNetworks.py:113-120."""
    o = F.relu(F.conv2d(x, sd[f"{q}.conv1.conv2d.weight"], sd[f"{q}.conv1.conv2d.bias"], padding=1))
    return F.conv2d(o, sd[f"{q}.conv2.conv2d.weight"], sd[f"{q}.conv2.conv2d.bias"], padding=1) * 0.1 + x


def _cf_up(sd: SD, q: str, x: torch.Tensor) -> torch.Tensor:
    """UpsampleConvLayer.forward (ConvTranspose2d k4 s2 p1), ChangeFormerBaseNetworks.py:96-105."""
    return F.conv_transpose2d(x, sd[f"{q}.conv2d.weight"], sd[f"{q}.conv2d.bias"], stride=2, padding=1)


def changeformer_v1_forward(sd: SD, x1: torch.Tensor, x2: torch.Tensor) -> torch.Tensor:
    """ChangeFormerV1.forward, ChangeFormer.py:657-674 with convprojection_base.forward (:605-641; the F.pad branches only
    fire for sizes that are not multiples of 32)."""
    f1 = mit_encoder_features(sd, x1, "Tenc", depths=(3, 4, 6, 3))
    f2 = mit_encoder_features(sd, x2, "Tenc", depths=(3, 4, 6, 3))
    di = [torch.abs(a - b) for a, b in zip(f1, f2)]
    c = "convproj"
    r = _cf_resblock(sd, f"{c}.dense_4.0", _cf_up(sd, f"{c}.convd16x", di[3])) + di[2]
    r = _cf_resblock(sd, f"{c}.dense_3.0", _cf_up(sd, f"{c}.convd8x", r)) + di[1]
    r = _cf_resblock(sd, f"{c}.dense_2.0", _cf_up(sd, f"{c}.convd4x", r)) + di[0]
    r = _cf_up(sd, f"{c}.convd1x", _cf_resblock(sd, f"{c}.dense_1.0", _cf_up(sd, f"{c}.convd2x", r)))
    return F.conv2d(r, sd["change_probability.conv2d.weight"], sd["change_probability.conv2d.bias"], padding=1)


def changeformer_v2_forward(sd: SD, x1: torch.Tensor, x2: torch.Tensor) -> torch.Tensor:
    """ChangeFormerV2.forward, ChangeFormer.py:931-948 with TDec.forward (:762-790)."""
    f1 = mit_encoder_features(sd, x1, "Tenc", depths=(3, 4, 6, 3))
    f2 = mit_encoder_features(sd, x2, "Tenc", depths=(3, 4, 6, 3))
    di = [torch.abs(a - b) for a, b in zip(f1, f2)]
    d = "TDec"
    size = di[0].shape[2:]
    cs = []
    for k in (4, 3, 2, 1):
        t = di[k - 1]
        n, _, hh, ww = t.shape
        y = F.linear(t.flatten(2).transpose(1, 2), sd[f"{d}.linear_c{k}.proj.weight"], sd[f"{d}.linear_c{k}.proj.bias"])
        y = y.permute(0, 2, 1).reshape(n, -1, hh, ww)
        cs.append(y if k == 1 else F.interpolate(y, size=size, mode="bilinear", align_corners=False))
    x = F.conv2d(torch.cat(cs, dim=1), sd[f"{d}.linear_fuse.weight"], sd[f"{d}.linear_fuse.bias"])
    x = _cf_resblock(sd, f"{d}.dense_2x.0", _cf_up(sd, f"{d}.convd2x", x))
    x = _cf_resblock(sd, f"{d}.dense_1x.0", _cf_up(sd, f"{d}.convd1x", x))
    return F.conv2d(x, sd[f"{d}.change_probability.conv2d.weight"], sd[f"{d}.change_probability.conv2d.bias"], padding=1)


def changeformer_v3_forward(sd: SD, x1: torch.Tensor, x2: torch.Tensor) -> torch.Tensor:
    """ChangeFormerV3.forward, ChangeFormer.py:966-973 with TDecV2.forward (:867-915): per-scale Linear heads on both dates, bilinear
    resize to the 1/4 scale, |.-.| per scale, 1x1 fuse, 3x3 conv to 16 * n_class channels + ReLU, PixelShuffle(4)."""
    f1 = mit_encoder_features(sd, x1, "Tenc", depths=(3, 4, 6, 3))
    f2 = mit_encoder_features(sd, x2, "Tenc", depths=(3, 4, 6, 3))
    d = "TDec"
    size = f1[0].shape[2:]

    def head(k, t):
        n, _, hh, ww = t.shape
        y = F.linear(t.flatten(2).transpose(1, 2), sd[f"{d}.linear_c{k}.proj.weight"], sd[f"{d}.linear_c{k}.proj.bias"])
        y = y.permute(0, 2, 1).reshape(n, -1, hh, ww)
        return y if k == 1 else F.interpolate(y, size=size, mode="bilinear", align_corners=False)

    diffs = [torch.abs(head(k, f1[k - 1]) - head(k, f2[k - 1])) for k in (4, 3, 2, 1)]
    c = F.conv2d(torch.cat(diffs, dim=1), sd[f"{d}.linear_fuse.weight"], sd[f"{d}.linear_fuse.bias"])
    x = F.relu(F.conv2d(c, sd[f"{d}.pix_shuffle_conv.weight"], sd[f"{d}.pix_shuffle_conv.bias"], padding=1))
    return F.pixel_shuffle(x, 4)
