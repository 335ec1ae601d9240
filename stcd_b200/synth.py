"""Seeded synthetic weights, image pairs and labels for parity tests and bench.py.

There is no network for datasets or checkpoints, so every measurement runs on synthetic data
of the reference's shapes: ``x1, x2 ~ randn(B, 3, H, W)`` like the reference's own smoke blocks
(models/SNUNet.py:248-249) and labels ~ Bernoulli(p) int64 {0,1} (data/dataset.py:206-210).

Weights: the module's constructor init, then BatchNorm running statistics and affine
parameters are randomised so that eval-mode BatchNorm is a non-trivial affine map and
activations stay O(1) through depth (with untouched running stats the logits
of a random-init net are degenerate and a parity test on them says nothing).  Everything is
drawn from a CPU ``torch.Generator`` so the same seed gives the same bits on every box.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn as nn

WEIGHT_SEED = 1337           # the reference's seed (train_stcd.py:62-65)
DATA_SEED = 1338

# Per-net weight gains of the parity / bench harness.  The north-star tolerance is ABSOLUTE (2e-2 on
# the logits) while the bf16 path's error is RELATIVE (~1 % of the activation scale after ~24 fused
# layers, 5-sigma tail over 1e5 logits), so the harness keeps the logit standard deviation near 0.2-0.3:
# non-degenerate change maps (tens of % "changed") with the tolerance still meaningful.
GAINS = {"SiamUnet_diff": 0.77, "SiamUnet_conc": 0.70, "SiamUnet_sub": 0.72, "SiamUnet_cross_conc": 0.78, "Unet": 0.64, "SNUNet_ECAM": 0.64, "SegCD": 0.70, "ChangeGNNV1": 0.5, "ChangeFormerV6": 0.5,
         "CDNet_model": 0.6, "BASE_Transformer": 0.47, "ResNet": 0.6, "DSIFN": 0.85, "ChangeGNNV2": 0.5, "ChangeGNNV2_Compare": 0.45, "VIG_V20_2": 0.47, "ChangeFormerV1": 0.38, "ChangeFormerV2": 0.5, "ChangeFormerV3": 0.42}
# Head-bias offsets (parameter name, per-class values added after the random draw) that centre the
# class margin of nets whose random-init margin is one-sided (SNUNet's post-ReLU features make class
# 0 win everywhere): without it the change map is all-zero and pixel agreement says nothing.
HEAD_BIAS = {"SNUNet_ECAM": ("conv_final.bias", [0.0, 0.95]), "CDNet_model": ("finalconv3_master.bias", [0.245, 0.0]),
             "BASE_Transformer": ("classifier.3.bias", [0.19, 0.0]), "DSIFN": ("o5_conv4.bias", [0.16]),
             "ChangeGNNV2_Compare": ("decoder.change_probability.conv2d.bias", [0.1, 0.0]),
             "ChangeFormerV1": ("change_probability.conv2d.bias", [0.2, 0.0])}


@torch.no_grad()
def randomize_(net: nn.Module, seed: int = WEIGHT_SEED, gain: float = 1.0) -> nn.Module:
    """Re-draw every parameter/buffer of `net` in a fixed order from a seeded CPU generator."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    for name, m in net.named_modules():
        if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d, nn.Linear)):
            w = m.weight
            if isinstance(m, nn.ConvTranspose2d):
                fan_in = w.shape[0] * w[0, 0].numel() / max(1, m.stride[0] * m.stride[1]) / m.groups
            elif isinstance(m, nn.Conv2d):
                fan_in = w.shape[1] * w[0, 0].numel()
            else:
                fan_in = w.shape[1]
            std = gain * (2.0 / max(1.0, fan_in)) ** 0.5
            w.copy_(torch.randn(w.shape, generator=g) * std)
            if m.bias is not None:
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
        elif isinstance(m, nn.BatchNorm2d):
            m.weight.copy_(0.8 + 0.4 * torch.rand(m.weight.shape, generator=g))
            m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
            m.running_var.copy_(0.5 + torch.rand(m.running_var.shape, generator=g))
        elif isinstance(m, nn.LayerNorm):
            m.weight.copy_(0.8 + 0.4 * torch.rand(m.weight.shape, generator=g))
            m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
        elif isinstance(m, nn.PReLU):
            m.weight.copy_(0.1 + 0.3 * torch.rand(m.weight.shape, generator=g))
        pe = m._parameters.get("pos_embed") if hasattr(m, "_parameters") else None
        if pe is not None:           # ViG / transformer positional embedding: zeros at construction (ChangeVIG.py:50)
            pe.copy_(torch.randn(pe.shape, generator=g) * 0.1)
        pe = m._parameters.get("pos_embedding") if hasattr(m, "_parameters") else None
        if pe is not None:           # BIT's learned token positions: global-RNG randn at construction (networks.py:337)
            pe.copy_(torch.randn(pe.shape, generator=g))
    return net


@torch.no_grad()
def prepare_(net: nn.Module, name: str, seed: int = WEIGHT_SEED) -> nn.Module:
    """The harness' weights for net family `name` (works on our wrappers and on the reference modules
    alike: same parameter names): seeded random draw at GAINS[name] + the HEAD_BIAS offset."""
    randomize_(net, seed=seed, gain=GAINS[name])
    if name in HEAD_BIAS:
        pname, vals = HEAD_BIAS[name]
        bias = dict(net.named_parameters())[pname]
        bias.add_(torch.tensor(vals[: bias.numel()], dtype=bias.dtype))
    if hasattr(net, "invalidate_plans"):
        net.invalidate_plans()
    return net


def image_pairs(batch: int, h: int, w: int, channels: int = 3, seed: int = DATA_SEED, correlated: float = 0.7
                ) -> Tuple[torch.Tensor, torch.Tensor]:
    """fp32 NCHW pairs.  T2 = correlated*T1 + noise: bi-temporal images are mostly unchanged, which
    keeps the change map non-degenerate (a few % .. tens of % of pixels 'changed')."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    x1 = torch.randn(batch, channels, h, w, generator=g)
    n = torch.randn(batch, channels, h, w, generator=g)
    x2 = correlated * x1 + (1.0 - correlated ** 2) ** 0.5 * n
    return x1, x2


def labels(batch: int, h: int, w: int, p: float = 0.05, seed: int = DATA_SEED + 1) -> torch.Tensor:
    """int64 {0,1} [B, H, W] ~ Bernoulli(p) (LEVIR-CD-like change sparsity)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.rand(batch, h, w, generator=g) < p).to(torch.int64)


def state_dict_cpu(net: nn.Module) -> Dict[str, torch.Tensor]:
    return {k: v.detach().to("cpu").clone() for k, v in net.state_dict().items()}
