"""A stand-in for the ViG ``gcn_lib`` package the reference imports but does not ship.  TEST INFRASTRUCTURE.

``models/pyramid_vig.py:17`` does ``from gcn_lib import Grapher, act_layer``; the package is absent from the reference
tree, not vendored and not version-pinned.  This module restates upstream ``vig_pytorch/gcn_lib`` (torch_vertex.py:
Grapher / DyGraphConv2d / MRConv2d; torch_nn.py: BasicConv / act_layer; pos_embed.py) on top of ``oracle/gcn.py`` so that
the reference's OWN ``ChangeVIG.py`` / ``pyramid_vig.py`` can be imported and run here (``refimport.install()`` puts it in
``sys.modules['gcn_lib']``).  **Parity unpinned**: the reference holds no test or fixture for it; module / parameter
names follow upstream (``fc1.0``, ``graph_conv.gconv.nn.0``, ``fc2.0``, ``relative_pos``).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from oracle import gcn


def act_layer(act: str, inplace: bool = False, neg_slope: float = 0.2, n_prelu: int = 1) -> nn.Module:
    act = act.lower()
    if act == "relu":
        return nn.ReLU(inplace)
    if act == "leakyrelu":
        return nn.LeakyReLU(neg_slope, inplace)
    if act == "prelu":
        return nn.PReLU(num_parameters=n_prelu, init=neg_slope)
    if act == "gelu":
        return nn.GELU()
    if act == "hswish":
        return nn.Hardswish(inplace)
    raise NotImplementedError("activation layer [%s] is not found" % act)


def _sincos_1d(embed_dim: int, pos: np.ndarray) -> np.ndarray:
    omega = np.arange(embed_dim // 2, dtype=np.float64) / (embed_dim / 2.0)
    omega = 1.0 / 10000 ** omega
    out = np.einsum("m,d->md", pos.reshape(-1), omega)
    return np.concatenate([np.sin(out), np.cos(out)], axis=1)


def get_2d_relative_pos_embed(embed_dim: int, grid_size: int) -> np.ndarray:
    """pos_embed.py: 2-D sin-cos embedding E [grid^2, embed_dim]; relative_pos = 2 E E^T / embed_dim."""
    gh = np.arange(grid_size, dtype=np.float32)
    gw = np.arange(grid_size, dtype=np.float32)
    grid = np.stack(np.meshgrid(gw, gh), axis=0).reshape([2, 1, grid_size, grid_size])
    emb = np.concatenate([_sincos_1d(embed_dim // 2, grid[0]), _sincos_1d(embed_dim // 2, grid[1])], axis=1)
    return 2 * np.matmul(emb, emb.transpose()) / emb.shape[1]


class BasicConv(nn.Sequential):
    def __init__(self, channels, act="relu", norm=None, bias=True):
        m = []
        for i in range(1, len(channels)):
            m.append(nn.Conv2d(channels[i - 1], channels[i], 1, bias=bias, groups=4))
            if norm is not None and norm.lower() != "none":
                m.append(nn.BatchNorm2d(channels[-1], affine=True))
            if act is not None and act.lower() != "none":
                m.append(act_layer(act))
        super().__init__(*m)


class MRConv2d(nn.Module):
    def __init__(self, in_channels, out_channels, act="relu", norm=None, bias=True):
        super().__init__()
        self.nn = BasicConv([in_channels * 2, out_channels], act, norm, bias)

    def forward(self, x, edge_index, y=None):
        return self.nn(gcn.mr_features(x, edge_index, y))


class DyGraphConv2d(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=9, dilation=1, conv="edge", act="relu", norm=None, bias=True,
                 stochastic=False, epsilon=0.0, r=1):
        super().__init__()
        if conv != "mr":
            raise NotImplementedError("only the max-relative conv the reference selects (ChangeVIG.py:309: conv='mr')")
        self.gconv = MRConv2d(in_channels, out_channels, act, norm, bias)
        self.k, self.d, self.r = kernel_size, dilation, r

    def forward(self, x, relative_pos=None):
        b, c, h, w = x.shape
        y = None
        if self.r > 1:
            y = F.avg_pool2d(x, self.r, self.r).reshape(b, c, -1, 1).contiguous()
        x = x.reshape(b, c, -1, 1).contiguous()
        with torch.no_grad():
            edge_index = gcn.dense_dilated_knn_graph(x, y, self.k, self.d, relative_pos)
        return self.gconv(x, edge_index, y).reshape(b, -1, h, w).contiguous()


class Grapher(nn.Module):
    def __init__(self, in_channels, kernel_size=9, dilation=1, conv="edge", act="relu", norm=None, bias=True, stochastic=False,
                 epsilon=0.0, r=1, n=196, drop_path=0.0, relative_pos=False):
        super().__init__()
        self.channels, self.n, self.r = in_channels, n, r
        self.fc1 = nn.Sequential(nn.Conv2d(in_channels, in_channels, 1, stride=1, padding=0), nn.BatchNorm2d(in_channels))
        self.graph_conv = DyGraphConv2d(in_channels, in_channels * 2, kernel_size, dilation, conv, act, norm, bias, stochastic,
                                        epsilon, r)
        self.fc2 = nn.Sequential(nn.Conv2d(in_channels * 2, in_channels, 1, stride=1, padding=0), nn.BatchNorm2d(in_channels))
        self.relative_pos = None
        if relative_pos:
            t = torch.from_numpy(np.float32(get_2d_relative_pos_embed(in_channels, int(n ** 0.5)))).unsqueeze(0).unsqueeze(1)
            t = F.interpolate(t, size=(n, n // (r * r)), mode="bicubic", align_corners=False)
            self.relative_pos = nn.Parameter(-t.squeeze(1), requires_grad=False)

    def _get_relative_pos(self, relative_pos, h, w):
        if relative_pos is None or h * w == self.n:
            return relative_pos
        n = h * w
        return F.interpolate(relative_pos.unsqueeze(0), size=(n, n // (self.r * self.r)), mode="bicubic").squeeze(0)

    def forward(self, x):
        tmp = x
        x = self.fc1(x)
        _, _, h, w = x.shape
        x = self.graph_conv(x, self._get_relative_pos(self.relative_pos, h, w))
        x = self.fc2(x)
        return x + tmp
