"""Lowering of reference modules to the fused-op program libstcd_b200 executes.

A *program* is a list of ops over NHWC bf16 tensors holding ``mult * chunk`` images (``mult`` = 2
for tensors that carry both temporal streams: T1 images first, then T2 images).  The only
compute op is the implicit-GEMM convolution (``ConvSpec``): its A operand is described by a
K-program (list of ``KEntry``: which source tensor, which filter tap, which channel block),
its B operand is a host-packed bf16 weight matrix, and its epilogue carries the folded
BatchNorm, ReLU, residual add, Siamese ``|f1 - f2|`` and 2x2 max-pool.

Everything here is pure host logic (numpy/torch on CPU): it is what ``tests/`` checks against
the reference without a GPU (through ``oracle/emulate.py``), and what ``plan.py`` hands to the
C-ABI on the GPU box.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

TILE_H, TILE_W = 8, 16
MAX_SRC = 6


def _bf16_bits(x: torch.Tensor) -> np.ndarray:
    """fp32 tensor -> uint16 bit patterns of its round-to-nearest-even bf16 value."""
    return x.detach().to(torch.float32).contiguous().to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)


def bf16_bits_to_f32(bits: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(bits.view(np.int16).copy()).view(torch.bfloat16).to(torch.float32)


@dataclass
class TensorSpec:
    name: str
    mult: int  # images = mult * chunk
    h: int
    w: int
    c: int


@dataclass
class KEntry:
    src: int      # index into ConvSpec.srcs
    dy: int       # input row offset (un-strided source pixels)
    dx: int
    c0: int       # first channel of the block inside the source tensor
    stream: int   # image offset in units of `chunk` (0: same image, 1: the T2 partner)
    wk: int       # first weight column


@dataclass
class Phase:
    k_begin: int
    k_count: int
    oy: int
    ox: int
    w_row: int


@dataclass
class ConvSpec:
    name: str
    srcs: List[str]
    src_sy: List[int]
    src_sx: List[int]
    hg: int
    wg: int
    img_mult: int
    pair: bool
    weights: np.ndarray          # uint16 [w_rows, w_cols] (bf16 bits), K-major
    kc: int
    n_tile: int
    cout: int
    cout_pad: int
    phases: List[Phase]
    kprog: List[KEntry]
    osy: int
    osx: int
    scale: np.ndarray            # float32 [cout_pad]
    shift: np.ndarray
    scale2: Optional[np.ndarray] = None
    shift2: Optional[np.ndarray] = None
    relu: bool = False
    res: Optional[str] = None
    out0: Optional[str] = None
    out0_coff: int = 0
    out_raw: Optional[str] = None
    out_pool: Optional[str] = None
    out_diff: Optional[str] = None
    out_ext: int = -1
    macs_per_pair: int = 0       # reference-equivalent MACs (for the roofline), per image pair


@dataclass
class InputPackSpec:
    name: str
    dst: str
    cin: int


@dataclass
class ExtOutput:
    name: str
    channels: int
    h: int
    w: int


@dataclass
class Program:
    model: str
    in_channels: int
    h: int
    w: int
    tensors: Dict[str, TensorSpec] = field(default_factory=dict)
    ops: List[object] = field(default_factory=list)
    ext: List[ExtOutput] = field(default_factory=list)

    def tensor(self, name: str, mult: int, h: int, w: int, c: int) -> str:
        if name in self.tensors:
            raise ValueError(f"duplicate tensor {name}")
        if c % 8:
            raise ValueError(f"tensor {name}: channels {c} must be a multiple of 8")
        self.tensors[name] = TensorSpec(name, mult, h, w, c)
        return name

    def macs_per_pair(self) -> int:
        return sum(getattr(o, "macs_per_pair", 0) for o in self.ops)


# ------------------------------------------------------------------------------------------
# folding


def fold_bn(bias: Optional[torch.Tensor], bn: Optional[Dict[str, torch.Tensor]], cout: int, eps: float = 1e-5
            ) -> Tuple[np.ndarray, np.ndarray]:
    """Eval-mode BatchNorm2d after a biased conv as one affine: y = acc*scale + shift.

    scale = gamma / sqrt(var + eps); shift = beta + (bias - mean) * scale  (SURVEY App. A;
    reference: SiamUnet_diff.py:99 ``relu(bn11(conv11(x)))`` with running statistics).
    """
    b = torch.zeros(cout, dtype=torch.float64) if bias is None else bias.detach().to(torch.float64)
    if bn is None:
        return np.ones(cout, np.float32), b.to(torch.float32).numpy()
    g = bn["weight"].detach().to(torch.float64)
    beta = bn["bias"].detach().to(torch.float64)
    mean = bn["running_mean"].detach().to(torch.float64)
    var = bn["running_var"].detach().to(torch.float64)
    scale = g / torch.sqrt(var + eps)
    shift = beta + (b - mean) * scale
    return scale.to(torch.float32).numpy(), shift.to(torch.float32).numpy()


def bn_params(sd: Dict[str, torch.Tensor], prefix: str) -> Dict[str, torch.Tensor]:
    return {k: sd[f"{prefix}.{k}"] for k in ("weight", "bias", "running_mean", "running_var")}


def _pad_vec(v: np.ndarray, n: int, fill: float) -> np.ndarray:
    out = np.full(n, fill, np.float32)
    out[: v.shape[0]] = v
    return out


# ------------------------------------------------------------------------------------------
# K-program / weight packing


@dataclass
class Segment:
    """A channel segment of the virtual concat: channels [0, c_real) of `tensor` (whose storage
    may be padded to more channels) at temporal stream `stream`."""
    tensor: str
    c_real: int
    stream: int = 0
    sy: int = 1
    sx: int = 1


def choose_kc(c_list: Sequence[int]) -> int:
    for kc in (64, 32, 16):
        if all(c % kc == 0 for c in c_list):
            return kc
    raise ValueError(f"channel counts {list(c_list)} are not multiples of 16")


def choose_n_tile(cout: int, pair: bool) -> Tuple[int, int]:
    cp = (cout + 15) // 16 * 16
    limit = 128 if pair else 256
    if cp <= limit:
        return cp, cp
    cp = (cout + 127) // 128 * 128
    return 128, cp


def _taps_to_gemm(
    prog: Program,
    name: str,
    segs: Sequence[Segment],
    phase_taps: Sequence[Tuple[int, int, List[Tuple[int, int, torch.Tensor]]]],
    cout: int,
    pair: bool,
) -> Tuple[np.ndarray, int, int, int, List[Phase], List[KEntry], List[str]]:
    """Build (weights, kc, n_tile, cout_pad, phases, kprog, srcs).

    phase_taps: per phase (oy, ox, [(dy, dx, W[cout, cin_total])...]) where dy/dx are offsets in
    tile-pixel units *before* the per-source stride is applied (the K entry stores
    dy*1, the kernel adds tile_origin*stride), and cin_total indexes the concatenated segments.
    """
    srcs: List[str] = []
    for s in segs:
        if s.tensor not in srcs:
            srcs.append(s.tensor)
    if len(srcs) > MAX_SRC:
        raise ValueError(f"{name}: {len(srcs)} sources > {MAX_SRC}")
    stored_c = [prog.tensors[s.tensor].c for s in segs]
    kc = choose_kc(stored_c)
    n_tile, cout_pad = choose_n_tile(cout, pair)
    phases: List[Phase] = []
    kprog: List[KEntry] = []
    cols: List[List[torch.Tensor]] = []
    for (oy, ox, taps) in phase_taps:
        k_begin = len(kprog)
        wk = 0
        blocks: List[torch.Tensor] = []
        for (dy, dx, wtap) in taps:
            ci = 0
            for s, sc in zip(segs, stored_c):
                wseg = torch.zeros(cout_pad, sc, dtype=torch.float32)
                wseg[:cout, : s.c_real] = wtap[:, ci: ci + s.c_real]
                ci += s.c_real
                for c0 in range(0, sc, kc):
                    blk = wseg[:, c0: c0 + kc]
                    if c0 >= s.c_real:  # pure padding block: contributes nothing, skip it
                        continue
                    kprog.append(KEntry(srcs.index(s.tensor), dy, dx, c0, s.stream, wk))
                    blocks.append(blk)
                    wk += kc
            if ci != wtap.shape[1]:
                raise ValueError(f"{name}: weight has {wtap.shape[1]} input channels, segments give {ci}")
        phases.append(Phase(k_begin, len(kprog) - k_begin, oy, ox, len(phases) * cout_pad))
        cols.append(blocks)
    w_cols = max(sum(b.shape[1] for b in blocks) for blocks in cols)
    wmat = torch.zeros(len(phases) * cout_pad, w_cols, dtype=torch.float32)
    for ph, blocks in enumerate(cols):
        if blocks:
            row = torch.cat(blocks, dim=1)
            wmat[ph * cout_pad: (ph + 1) * cout_pad, : row.shape[1]] = row
    return _bf16_bits(wmat), kc, n_tile, cout_pad, phases, kprog, srcs


def conv_taps(weight: torch.Tensor, pad: int, stride: int = 1) -> List[Tuple[int, int, List]]:
    """nn.Conv2d weight [cout, cin, kh, kw] -> one phase of taps (dy, dx, W[cout, cin])."""
    kh, kw = weight.shape[2], weight.shape[3]
    taps = [(ky - pad, kx - pad, weight[:, :, ky, kx].to(torch.float32)) for ky in range(kh) for kx in range(kw)]
    return [(0, 0, taps)]


def convT_as_conv_weight(weight_t: torch.Tensor) -> torch.Tensor:
    """ConvTranspose2d(stride=1) weight [cin, cout, k, k] -> equivalent Conv2d weight
    [cout, cin, k, k] (flip both spatial axes, swap channel axes); padding' = k - 1 - padding.
    Reference: the FC-Siam decoder "convs" are ConvTranspose2d(k=3, padding=1): SiamUnet_diff.py:54-90."""
    return weight_t.flip(2, 3).transpose(0, 1).contiguous()


def convT_phase_taps(weight_t: torch.Tensor, stride: int, pad: int) -> List[Tuple[int, int, List]]:
    """ConvTranspose2d weight [cin, cout, k, k], stride s -> s*s output phases.

    out[s*i + a] = sum over ky with (a + pad - ky) % s == 0 of in[i + (a + pad - ky)//s] * W[ky]
    (SiamUnet_diff.py:52 k3 s2 p1 op1; SNUNet.py:38 k2 s2; ChangeFormerBaseNetworks.py:101 k4 s2 p1).
    """
    k = weight_t.shape[2]
    out = []

    def axis(a):
        return [((a + pad - kk) // stride, kk) for kk in range(k) if (a + pad - kk) % stride == 0]

    for a in range(stride):
        for b in range(stride):
            taps = [(dy, dx, weight_t[:, :, ky, kx].transpose(0, 1).to(torch.float32))
                    for (dy, ky) in axis(a) for (dx, kx) in axis(b)]
            out.append((a, b, taps))
    return out


def add_conv(
    prog: Program,
    name: str,
    segs: Sequence[Segment],
    phase_taps,
    cout: int,
    hg: int,
    wg: int,
    img_mult: int,
    scale: np.ndarray,
    shift: np.ndarray,
    *,
    pair: bool = False,
    relu: bool = False,
    osy: int = 1,
    osx: int = 1,
    scale2: Optional[np.ndarray] = None,
    shift2: Optional[np.ndarray] = None,
    res: Optional[str] = None,
    out0: Optional[str] = None,
    out0_coff: int = 0,
    out_raw: Optional[str] = None,
    out_pool: Optional[str] = None,
    out_diff: Optional[str] = None,
    out_ext: int = -1,
    macs_per_pair: int = 0,
) -> ConvSpec:
    wbits, kc, n_tile, cout_pad, phases, kprog, srcs = _taps_to_gemm(prog, name, segs, phase_taps, cout, pair)
    sy = [1] * len(srcs)
    sx = [1] * len(srcs)
    for s in segs:
        sy[srcs.index(s.tensor)] = s.sy
        sx[srcs.index(s.tensor)] = s.sx
    # K entries carry offsets in un-strided source pixels
    spec = ConvSpec(
        name=name, srcs=srcs, src_sy=sy, src_sx=sx, hg=hg, wg=wg, img_mult=img_mult, pair=pair,
        weights=wbits, kc=kc, n_tile=n_tile, cout=cout, cout_pad=cout_pad, phases=phases, kprog=kprog,
        osy=osy, osx=osx, scale=_pad_vec(scale, cout_pad, 1.0), shift=_pad_vec(shift, cout_pad, 0.0),
        scale2=None if scale2 is None else _pad_vec(scale2, cout_pad, 1.0),
        shift2=None if shift2 is None else _pad_vec(shift2, cout_pad, 0.0),
        relu=relu, res=res, out0=out0, out0_coff=out0_coff, out_raw=out_raw, out_pool=out_pool,
        out_diff=out_diff, out_ext=out_ext, macs_per_pair=macs_per_pair,
    )
    prog.ops.append(spec)
    return spec
