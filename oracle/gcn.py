"""CPU restatement of the ViG ``gcn_lib`` graph ops the reference imports but does not ship.  TEST INFRASTRUCTURE.

**Parity unpinned**: ``models/pyramid_vig.py:17`` does ``from gcn_lib import Grapher, act_layer`` and the module is
absent from the reference tree, not vendored and not version-pinned (no requirements file).  Upstream is
huawei-noah/Efficient-AI-Backbones, ``vig_pytorch/gcn_lib/{torch_edge,torch_vertex}.py``; this file restates its
published algorithm, anchored on the reference's own call sites for the constructor
arguments (``ChangeVIG.py:61-63``: kernel_size k=9, dilation min(idx//4+1, 5), conv='mr', r = reduce ratio,
relative_pos=True).  The reference holds no test or golden vector for these ops.

Layouts are the reference's: node features fp32 ``[B, C, N, 1]`` (N = H*W nodes), neighbour tables int64.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F


def pairwise_distance(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """torch_edge.xy_pairwise_distance: x [B, N, C], y [B, M, C] -> |x_i|^2 - 2 x_i.y_j + |y_j|^2, [B, N, M]."""
    xy_inner = -2 * torch.matmul(x, y.transpose(2, 1))
    x_square = torch.sum(torch.mul(x, x), dim=-1, keepdim=True)
    y_square = torch.sum(torch.mul(y, y), dim=-1, keepdim=True)
    return x_square + xy_inner + y_square.transpose(2, 1)


def dense_dilated_knn_graph(x: torch.Tensor, y: Optional[torch.Tensor], k: int, dilation: int,
                            relative_pos: Optional[torch.Tensor] = None) -> torch.Tensor:
    """torch_edge.DenseDilatedKnnGraph.forward (eval mode: stochastic dilation is training-only).

    x [B, C, N, 1]; y [B, C, M, 1] or None (then y := x); relative_pos [1, N, M] or None.
    Nodes are L2-normalised over channels, the k*dilation nearest y-nodes of every x-node are found
    (ascending distance), and every dilation-th of them is kept.
    Returns edge_index int64 [2, B, N, k]: [0] = neighbour index (into y), [1] = centre index."""
    x = F.normalize(x, p=2.0, dim=1)
    y = x if y is None else F.normalize(y, p=2.0, dim=1)
    xn = x.transpose(2, 1).squeeze(-1)                     # [B, N, C]
    yn = y.transpose(2, 1).squeeze(-1)
    b, n, _ = xn.shape
    dist = pairwise_distance(xn, yn)
    if relative_pos is not None:
        dist = dist + relative_pos
    _, nn_idx = torch.topk(-dist, k=k * dilation)          # [B, N, k*d]
    center_idx = torch.arange(0, n).repeat(b, k * dilation, 1).transpose(2, 1)
    edge_index = torch.stack((nn_idx, center_idx), dim=0)
    return edge_index[:, :, :, ::dilation]


def batched_index_select(x: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """torch_nn.batched_index_select: x [B, C, M, 1], idx [B, N, k] -> [B, C, N, k]."""
    b, c, m = x.shape[:3]
    _, n, k = idx.shape
    flat = (idx + torch.arange(0, b).view(-1, 1, 1) * m).contiguous().view(-1)
    feat = x.transpose(2, 1).contiguous().view(b * m, -1)[flat]
    return feat.view(b, n, k, c).permute(0, 3, 1, 2).contiguous()


def max_relative(x: torch.Tensor, edge_index: torch.Tensor, y: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The aggregation of torch_vertex.MRConv2d.forward: max_k (x_j - x_i), [B, C, N, 1]."""
    x_i = batched_index_select(x, edge_index[1])
    x_j = batched_index_select(x if y is None else y, edge_index[0])
    m, _ = torch.max(x_j - x_i, -1, keepdim=True)
    return m


def mr_features(x: torch.Tensor, edge_index: torch.Tensor, y: Optional[torch.Tensor] = None) -> torch.Tensor:
    """MRConv2d's input to its grouped 1x1 conv: channel-INTERLEAVED (x0, m0, x1, m1, ...), [B, 2C, N, 1]."""
    b, c, n, _ = x.shape
    m = max_relative(x, edge_index, y)
    return torch.cat([x.unsqueeze(2), m.unsqueeze(2)], dim=2).reshape(b, 2 * c, n, 1)
