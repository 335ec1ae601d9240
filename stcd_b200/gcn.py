"""ViG Grapher graph ops behind gcn_lib's call signatures (the ChangeVIG path, models/pyramid_vig.py:17).

``DenseDilatedKnnGraph(k, dilation)(x, y, relative_pos)`` and the max-relative aggregation of ``MRConv2d`` run
as hand-written CUDA kernels (csrc/graph_kernels.cuh) on the reference's own layouts (fp32 ``[B, C, N, 1]`` node
features, int64 ``[2, B, N, k]`` edge index).  No CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib


def _stream(t: torch.Tensor):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _nodes(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError("stcd_b200 has no CPU path: move the tensors to a B200 (cuda) device")
    if t.dtype != torch.float32 or t.dim() != 4 or t.shape[3] != 1:
        raise TypeError(f"{name} must be float32 [B, C, N, 1] (gcn_lib's node layout), got {t.dtype} {tuple(t.shape)}")
    return t.contiguous()


class DenseDilatedKnnGraph(torch.nn.Module):
    """gcn_lib.torch_edge.DenseDilatedKnnGraph (eval mode): returns edge_index int64 [2, B, N, k]."""

    def __init__(self, k: int = 9, dilation: int = 1, stochastic: bool = False, epsilon: float = 0.0):
        super().__init__()
        self.k, self.dilation = int(k), int(dilation)
        self.stochastic, self.epsilon = stochastic, epsilon     # training-time only upstream

    @torch.no_grad()
    def forward(self, x: torch.Tensor, y: Optional[torch.Tensor] = None, relative_pos: Optional[torch.Tensor] = None):
        x = _nodes(x, "x")
        b, c, n, _ = x.shape
        m = n
        if y is not None:
            y = _nodes(y, "y")
            m = y.shape[2]
            if y.shape[0] != b or y.shape[1] != c:
                raise ValueError(f"x {tuple(x.shape)} and y {tuple(y.shape)} disagree")
        rp = None
        if relative_pos is not None:
            rp = relative_pos.to(torch.float32).contiguous()
            if rp.numel() != n * m:
                raise ValueError(f"relative_pos has {rp.numel()} elements, expected [1, {n}, {m}]")
        nn_idx = torch.empty(b, n, self.k, dtype=torch.int64, device=x.device)
        scratch = torch.empty(b * c * (n + (m if y is not None else 0)), dtype=torch.float32, device=x.device)
        lib = _lib.lib()
        _lib.check(lib.stcd_knn_graph(x.data_ptr(), y.data_ptr() if y is not None else None,
                                      rp.data_ptr() if rp is not None else None, b, c, n, m, self.k, self.dilation,
                                      nn_idx.data_ptr(), scratch.data_ptr(), _stream(x)), "stcd_knn_graph")
        center = torch.arange(n, device=x.device).view(1, n, 1).expand(b, n, self.k)
        return torch.stack((nn_idx, center), dim=0)


@torch.no_grad()
def max_relative(x: torch.Tensor, edge_index: torch.Tensor, y: Optional[torch.Tensor] = None, interleave: bool = False):
    """max_k (x_j - x_i) of MRConv2d.forward: [B, C, N, 1]; ``interleave=True`` returns MRConv2d's conv input
    [B, 2C, N, 1] with channels (x0, m0, x1, m1, ...)."""
    x = _nodes(x, "x")
    b, c, n, _ = x.shape
    m = n
    if y is not None:
        y = _nodes(y, "y")
        m = y.shape[2]
    idx = edge_index[0].to(torch.int64).contiguous()
    if tuple(idx.shape[:2]) != (b, n):
        raise ValueError(f"edge_index[0] is {tuple(idx.shape)}, expected [{b}, {n}, k]")
    k = idx.shape[2]
    out = torch.empty(b, 2 * c if interleave else c, n, 1, dtype=torch.float32, device=x.device)
    lib = _lib.lib()
    _lib.check(lib.stcd_max_relative(x.data_ptr(), y.data_ptr() if y is not None else None, idx.data_ptr(), b, c, n, m, k,
                                     1 if interleave else 0, out.data_ptr(), _stream(x)), "stcd_max_relative")
    return out
