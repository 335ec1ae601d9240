"""CPU emulator of a lowered Program (stcd_b200/lowering.py).  TEST INFRASTRUCTURE.

Executes the op list exactly as csrc/conv_ws.cuh does — logical NHWC tensors holding bf16 values,
K-programs walked entry by entry, fp32 accumulation, the epilogue's rounding points — so that
``tests/`` can check the HOST-side lowering (weight packing, K-programs, BN folding, phases,
virtual concat) against the reference without a GPU, and the GPU kernels against this emulator
to fp32-accumulation-order accuracy.
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np
import torch

from stcd_b200 import lowering as L


def _bf16(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32)


def _store(T: Dict[str, torch.Tensor], name: str, idx, coff: int, v: torch.Tensor, split: bool) -> None:
    """Store logical channels v[..., :c] at channel offset `coff` of tensor `name`: one bf16 value per element, or (split
    precision) hi = bf16(v) in the hi plane and lo = bf16(v - hi) in the lo plane (second half of the stored channels)."""
    c = v.shape[-1]
    hi = _bf16(v)
    T[name][idx + (slice(coff, coff + c),)] = hi
    if split:
        half = T[name].shape[-1] // 2
        T[name][idx + (slice(half + coff, half + coff + c),)] = _bf16(v - hi)


def _load(T: Dict[str, torch.Tensor], name: str, idx, c: int, split: bool) -> torch.Tensor:
    t = T[name]
    v = t[idx + (slice(0, c),)]
    if split:
        half = t.shape[-1] // 2
        v = v + t[idx + (slice(half, half + c),)]
    return v


def _gather(src: torch.Tensor, rows: torch.Tensor, cols: torch.Tensor) -> torch.Tensor:
    """src [N,h,w,c]; rows [hg], cols [wg] (may be out of range -> zeros, like TMA OOB fill)."""
    h, w = src.shape[1], src.shape[2]
    rv = (rows >= 0) & (rows < h)
    cv = (cols >= 0) & (cols < w)
    g = src[:, rows.clamp(0, h - 1)][:, :, cols.clamp(0, w - 1)]
    return g * (rv[:, None] & cv[None, :]).to(src.dtype)[None, :, :, None]


def run_conv(op: L.ConvSpec, T: Dict[str, torch.Tensor], chunk: int, ext: List[torch.Tensor], n_valid: int) -> None:
    n_img = chunk if op.pair else op.img_mult * chunk
    n_m = 2 if op.pair else 1
    n_nt = op.cout_pad // op.n_tile
    ho, wo = op.hg * op.osy, op.wg * op.osx
    scale = torch.from_numpy(op.scale)
    shift = torch.from_numpy(op.shift)
    ii = torch.arange(op.hg)
    jj = torch.arange(op.wg)
    full = [torch.zeros(n_img, ho, wo, op.cout_pad) for _ in range(n_m)]
    for ph in op.phases:
        for m in range(n_m):
            acc = torch.zeros(n_img, op.hg, op.wg, op.cout_pad)
            blk = ph.w_block
            for ck in op.chunks[ph.chunk_begin: ph.chunk_begin + ph.chunk_count]:
                si = ck.src
                src = T[op.srcs[si]]
                off = (ck.stream + m) * chunk
                sub = src[off: off + n_img, :, :, ck.c0: ck.c0 + op.kc]
                if sub.shape[3] < op.kc:      # chunk overhangs the tensor's channels: TMA zero-fill
                    sub = torch.nn.functional.pad(sub, (0, op.kc - sub.shape[3]))
                for (ty, tx) in op.taps[ck.tap_begin: ck.tap_begin + ck.n_taps]:
                    a = _gather(sub, ii * op.src_sy[si] + ck.by + ty * op.src_sy[si],
                                jj * op.src_sx[si] + ck.bx + tx * op.src_sx[si])
                    w = torch.cat([op.weight_block(nt, blk) for nt in range(n_nt)], 0)      # [cout_pad, kc]
                    acc += a @ w.T
                    blk += 1
            assert blk == ph.w_block + ph.n_blocks
            if op.fold_cs:
                # columns = (folded phase, channel): affine (+ activation) per column, then column block p of GEMM phase
                # (oy, ox) -> output pixel phase (oy + p // osx, ox + p % osx)
                v = acc * scale + shift
                v = torch.relu(v) if op.act_kind == 1 else (torch.nn.functional.gelu(v) if op.act_kind == 2 else
                                                             (torch.where(v >= 0, v, v * op.act_alpha) if op.act_kind == 3 else v))
                sl = slice(m * chunk, m * chunk + n_img)
                for p_ in range(op.cout // op.fold_cs):
                    q = _bf16(v[..., p_ * op.fold_cs: p_ * op.fold_cs + op.fold_cout])
                    T[op.out0][sl, ph.oy + p_ // op.osx::op.osy, ph.ox + p_ % op.osx::op.osx,
                               op.out0_coff: op.out0_coff + op.fold_cout] = q
            elif op.xf_cs:
                # horizontal tap folding: column block b at input position x is tap (dy, b - 1)'s contribution to output
                # x - (b - 1);  out[x] = D_0[x - 1] + D_1[x] + D_2[x + 1]  (positions outside the image contribute zeros)
                cs = op.xf_cs
                d0, d1, d2 = acc[..., :cs], acc[..., cs: 2 * cs], acc[..., 2 * cs: 3 * cs]
                out = d1.clone()
                out[:, :, 1:] += d0[:, :, :-1]
                out[:, :, :-1] += d2[:, :, 1:]
                full[m][..., :cs] = out
            else:
                full[m][:, ph.oy::op.osy, ph.ox::op.osx] = acc
    def act(v):
        if op.act_kind == 1:
            return torch.relu(v)
        if op.act_kind == 2:
            return torch.nn.functional.gelu(v)
        if op.act_kind == 3:
            return torch.where(v >= 0, v, v * op.act_alpha)
        return v

    if op.fold_cs:
        return
    vs = []
    sp = bool(getattr(op, "split", False))
    for m in range(n_m):
        v = full[m] * scale + shift
        sl = slice(m * chunk, m * chunk + n_img)
        every = (sl, slice(None), slice(None))
        if op.out_raw is not None:
            _store(T, op.out_raw, every, 0, v[..., : op.cout], sp)
        if op.scale2 is not None:
            if op.act_pre:
                v = act(v)
            v = v * torch.from_numpy(op.scale2) + torch.from_numpy(op.shift2)
        if op.res is not None:
            v[..., : op.cout] = v[..., : op.cout] + _load(T, op.res, every, op.cout, sp)
        if not op.act_pre:
            v = act(v)
        if op.out0 is not None and op.out0_s2d:
            # space-to-depth store: pixel (y, x), channel c -> pixel (y//2, x//2), channel ((y%2)*2 + x%2)*cout + c
            q = _bf16(v[..., : op.cout])
            for py in range(2):
                for px in range(2):
                    k = (py * 2 + px) * op.cout
                    T[op.out0][sl, :, :, k: k + op.cout] = q[:, py::2, px::2]
        elif op.out0 is not None:
            _store(T, op.out0, every, op.out0_coff, v[..., : op.cout], sp)
        if op.out_ext >= 0:
            o = v[:n_valid, :, :, : op.cout].permute(0, 3, 1, 2)
            if len(op.phases) == op.osy * op.osx:
                ext[op.out_ext][:n_valid] = o
            else:                      # an op that owns only some output phases (PixelShuffle(4) as 4 ops x 4 phases) writes only those
                for ph in op.phases:
                    ext[op.out_ext][:n_valid, :, ph.oy::op.osy, ph.ox::op.osx] = o[:, :, ph.oy::op.osy, ph.ox::op.osx]
        if op.out_pool is not None:
            q = torch.maximum(torch.maximum(v[:, 0::2, 0::2], v[:, 0::2, 1::2]),
                              torch.maximum(v[:, 1::2, 0::2], v[:, 1::2, 1::2]))
            _store(T, op.out_pool, every, 0, q[..., : op.cout], sp)
        vs.append(v)
    if op.out_diff is not None:
        _store(T, op.out_diff, (slice(0, n_img), slice(None), slice(None)), 0, (vs[0] - vs[1]).abs()[..., : op.cout], sp)


def normalize_u8(img: torch.Tensor, mean, std) -> torch.Tensor:
    """The reference loader on a decoded image batch: uint8 HWC [B, H, W, C] -> fp32 NCHW, torchvision's
    ToTensor (x.div(255)) then Normalize (x.sub(mean).div(std)) in fp32 (data/dataset.py:196-203)."""
    x = img.permute(0, 3, 1, 2).to(torch.float32).div(255)
    m = torch.tensor(mean, dtype=torch.float32).view(1, -1, 1, 1)
    s = torch.tensor(std, dtype=torch.float32).view(1, -1, 1, 1)
    return x.sub(m).div(s)


def run_program(prog: L.Program, x1: torch.Tensor, x2: torch.Tensor, chunk: int | None = None,
                keep: Dict[str, torch.Tensor] | None = None) -> List[torch.Tensor]:
    """x1, x2: fp32 NCHW [B, cin, H, W] (or uint8 HWC when the input op carries u8_norm) -> list of fp32 NCHW
    external outputs."""
    for op in prog.ops:
        if isinstance(op, L.InputPackSpec) and op.u8_norm is not None:
            x1, x2 = normalize_u8(x1, *op.u8_norm), normalize_u8(x2, *op.u8_norm)
    n = x1.shape[0]
    chunk = chunk or n
    outs = [torch.zeros(n, e.channels, e.h, e.w) for e in prog.ext]
    for start in range(0, n, chunk):
        nv = min(chunk, n - start)
        T = {name: torch.zeros(t.mult * chunk, t.h, t.w, t.c) for name, t in prog.tensors.items()}
        for name, val in prog.consts.items():
            T[name][..., : val.shape[2]] = _bf16(val)[None]
        ext = [torch.zeros(chunk, e.channels, e.h, e.w) for e in prog.ext]
        for op in prog.ops:
            if isinstance(op, L.InputPackSpec) and op.s2d:
                dst = T[op.dst]
                for s_, xs in enumerate((x1, x2)):
                    v = _bf16(xs[start: start + nv].permute(0, 2, 3, 1))
                    for py in range(2):
                        for px in range(2):
                            k = (py * 2 + px) * op.cin
                            dst[s_ * chunk: s_ * chunk + nv, :, :, k: k + op.cin] = v[:, py::2, px::2]
            elif isinstance(op, L.InputPackSpec):
                every = (slice(None), slice(None))
                _store(T, op.dst, (slice(0, nv),) + every, 0, x1[start: start + nv].permute(0, 2, 3, 1), op.split)
                _store(T, op.dst, (slice(chunk, chunk + nv),) + every, 0, x2[start: start + nv].permute(0, 2, 3, 1), op.split)
            elif isinstance(op, L.ConvSpec):
                run_conv(op, T, chunk, ext, nv)
            else:
                run_aux(op, T, chunk, ext, nv)
        for k in range(len(outs)):
            outs[k][start: start + nv] = ext[k][:nv]
        if keep is not None:
            keep.update(T)
    return outs


def run_ecam_head(op: L.EcamHeadSpec, T: Dict[str, torch.Tensor], chunk: int, ext: List[torch.Tensor], nv: int) -> None:
    """csrc/aux_kernels.cuh ecam_stats_kernel + ecam_head_kernel: fp32 reductions over the bf16 level-0
    tensors, fp32 MLPs, then a per-image 1x1 head (models/SNUNet.py:144-149)."""
    xs = [_load(T, s, (slice(0, chunk), slice(None), slice(None)), op.c, op.split) for s in op.srcs]      # [chunk, h, w, C]
    out = torch.cat(xs, 3)                                  # [chunk, h, w, 4C]
    intra = xs[0] + xs[1] + xs[2] + xs[3]

    def attention(t, fc1, fc2):
        avg, mx = t.mean(dim=(1, 2)), t.amax(dim=(1, 2))    # [chunk, C']
        mlp = lambda v: torch.relu(v @ torch.from_numpy(fc1).T) @ torch.from_numpy(fc2).T  # noqa: E731
        return torch.sigmoid(mlp(avg) + mlp(mx))

    ca = attention(out, op.ca_fc1, op.ca_fc2)               # [chunk, 4C]
    ca1 = attention(intra, op.ca1_fc1, op.ca1_fc2)          # [chunk, C]
    wf = torch.from_numpy(op.w_final)                       # [k, 4C]
    w_eff = wf[None] * ca[:, None, :]                       # [chunk, k, 4C]
    b_eff = torch.from_numpy(op.b_final)[None] + (w_eff * ca1.repeat(1, 4)[:, None, :]).sum(2)
    y = torch.einsum("nhwc,nkc->nkhw", out, w_eff) + b_eff[:, :, None, None]
    ext[op.out_ext][:nv] = y[:nv]


def _from_s2d(t: torch.Tensor, c: int) -> torch.Tensor:
    """[n, h/2, w/2, 4c] space-to-depth -> [n, h, w, c]."""
    n, h2, w2, _ = t.shape
    full = torch.zeros(n, 2 * h2, 2 * w2, c)
    for py in range(2):
        for px in range(2):
            k = (py * 2 + px) * c
            full[:, py::2, px::2] = t[..., k: k + c]
    return full


def run_maxpool_s2d(op: L.MaxPoolS2DSpec, T: Dict[str, torch.Tensor]) -> None:
    """csrc/aux_kernels.cuh maxpool3x3s2_s2d_kernel: max over bf16 values is exact."""
    full = _from_s2d(T[op.src], op.c).permute(0, 3, 1, 2)
    T[op.dst][..., : op.c] = torch.nn.functional.max_pool2d(full, 3, 2, 1).permute(0, 2, 3, 1)


def run_seg_head(op: L.SegHeadSpec, T: Dict[str, torch.Tensor], chunk: int, ext: List[torch.Tensor], nv: int) -> None:
    """csrc/aux_kernels.cuh segcd_head_kernel: fp32 3x3 conv (fp32 weights) over the bf16 decoder output of
    both streams and over |d1 - d2| (exact in fp32), then change = min(head(|d1 - d2|), |m1 - m2|)."""
    d = T[op.src][..., : op.c].permute(0, 3, 1, 2)
    d1, d2 = d[:chunk], d[chunk: 2 * chunk]
    w = torch.from_numpy(op.weight).reshape(3, 3, op.c).permute(2, 0, 1)[None]      # [1, c, 3, 3]
    mma = op.c == 16 and op.diff_src is None      # segcd_head_mma_kernel: bf16 weights, |d1 - d2| rounded to bf16
    if mma:
        w = _bf16(w)
    b = torch.tensor([op.bias])
    head = lambda t: torch.nn.functional.conv2d(t, w, b, padding=1)  # noqa: E731
    m1, m2 = head(d1), head(d2)
    dd = (d1 - d2).abs() if op.diff_src is None else T[op.diff_src][:chunk, :, :, : op.c].permute(0, 3, 1, 2)
    if mma:
        dd = _bf16(dd)
    change = torch.minimum(head(dd), (m1 - m2).abs())
    ext[op.out_ext][:nv] = m1[:nv]
    ext[op.out_ext + 1][:nv] = m2[:nv]
    ext[op.out_ext + 2][:nv] = change[:nv]


def run_graph_conv(op: L.GraphConvSpec, T: Dict[str, torch.Tensor]) -> None:
    """csrc/graph_kernels.cuh on a bf16 tensor: fp32 kNN graph + max-relative, rounded to bf16 on store."""
    from oracle import gcn
    x = T[op.src][..., : op.c].permute(0, 3, 1, 2).contiguous()                  # [imgs, c, h, w]
    b, c, h, w = x.shape
    y = torch.nn.functional.avg_pool2d(x, op.r, op.r).reshape(b, c, -1, 1) if op.r > 1 else None
    xn = x.reshape(b, c, -1, 1)
    rp = None if op.relpos is None else torch.from_numpy(op.relpos)[None]
    e = gcn.dense_dilated_knn_graph(xn, y, op.k, op.dilation, rp)
    m = gcn.max_relative(xn, e, y).reshape(b, c, h, w)
    T[op.dst][..., : op.c] = _bf16(m.permute(0, 2, 3, 1))


def run_bilinear_up(op: L.BilinearUpSpec, T: Dict[str, torch.Tensor]) -> None:
    x = T[op.src][..., : op.c].permute(0, 3, 1, 2)
    y = torch.nn.functional.interpolate(x, scale_factor=op.scale, mode="bilinear", align_corners=False)
    T[op.dst][..., : op.c] = _bf16(y.permute(0, 2, 3, 1))


def run_layernorm(op: L.LayerNormSpec, T: Dict[str, torch.Tensor]) -> None:
    x = T[op.src][..., : op.c]
    y = _bf16(torch.nn.functional.layer_norm(x, (op.c,), torch.from_numpy(op.gamma), torch.from_numpy(op.beta), op.eps))
    T[op.dst][..., : op.c] = y
    if op.dst_s2d is not None:
        for py in range(2):
            for px in range(2):
                k = (py * 2 + px) * op.c
                T[op.dst_s2d][..., k: k + op.c] = y[:, py::2, px::2]


def run_attention(op: L.AttentionSpec, T: Dict[str, torch.Tensor]) -> None:
    q = T[op.q][..., : op.c]
    kv = T[op.kv][..., : 2 * op.c]
    b, h, w, c = q.shape
    d = c // op.heads
    qh = q.reshape(b, h * w, op.heads, d).permute(0, 2, 1, 3)
    k = kv[..., :c].reshape(b, -1, op.heads, d).permute(0, 2, 1, 3)
    v = kv[..., c:].reshape(b, -1, op.heads, d).permute(0, 2, 1, 3)
    s_ = (qh @ k.transpose(-2, -1)) * op.scale
    if k.shape[2] == 64:
        # the tensor-core kernel's rounding points: un-normalised probabilities exp(s - max) rounded to bf16 for P V, fp32 row sum
        p_ = torch.exp(s_ - s_.amax(dim=-1, keepdim=True))
        o = (_bf16(p_) @ v) / p_.sum(dim=-1, keepdim=True)
    else:
        o = s_.softmax(dim=-1) @ v
    T[op.dst][..., : c] = _bf16(o.transpose(1, 2).reshape(b, h, w, c))


def run_dwconv(op: L.DWConvSpec, T: Dict[str, torch.Tensor]) -> None:
    x = T[op.src][..., : op.c].permute(0, 3, 1, 2)
    wt = torch.from_numpy(op.weight).reshape(op.c, 1, 3, 3)
    y = torch.nn.functional.conv2d(x, wt, torch.from_numpy(op.bias), padding=1, groups=op.c)
    if op.gelu:
        y = torch.nn.functional.gelu(y)
    T[op.dst][..., : op.c] = _bf16(y.permute(0, 2, 3, 1))


def run_aux(op, T, chunk, ext, nv):
    if isinstance(op, L.AbsDiffSpec):
        d = T[op.src][:chunk, :, :, : op.c] - T[op.src][chunk: 2 * chunk, :, :, : op.c]
        if op.signed and op.add is not None:
            d = (T[op.add][..., : op.c] + T[op.src][:chunk, :, :, : op.c]) - T[op.src][chunk: 2 * chunk, :, :, : op.c]
        T[op.dst][..., : op.c] = _bf16(d if op.signed else d.abs())
        return None
    if isinstance(op, L.BitTransformerSpec):
        run_bit_transformer(op, T, chunk)
        return None
    if isinstance(op, L.ChannelAttentionSpec):
        x = torch.cat([T[t][s_ * chunk: (s_ + 1) * chunk, :, :, :c_] for (t, s_, c_) in op.srcs], dim=-1)      # [chunk, h, w, C]
        fc1, fc2 = torch.from_numpy(op.fc1), torch.from_numpy(op.fc2)
        hid = torch.relu(x.mean(dim=(1, 2)) @ fc1.T) + torch.relu(x.amax(dim=(1, 2)) @ fc1.T)
        T[op.dst][...] = _bf16(x * torch.sigmoid(hid @ fc2.T)[:, None, None, :])
        return None
    if isinstance(op, L.SpatialGateSpec):
        import torch.nn.functional as F
        x = T[op.src][..., : op.c]
        m = torch.stack([x.mean(dim=-1), x.amax(dim=-1)], dim=1)                                   # [n, 2, h, w]
        g = torch.sigmoid(F.conv2d(m, torch.from_numpy(op.w)[None], None, padding=3))[:, 0, :, :, None]
        T[op.dst][..., : op.c] = _bf16(x * g * torch.from_numpy(op.scale) + torch.from_numpy(op.shift))
        return None
    if isinstance(op, L.GlobalLocalGateSpec):
        import torch.nn.functional as F
        x = T[op.src][..., : op.c]
        ch = torch.relu((x.mean(dim=(1, 2)) * torch.from_numpy(op.w_avg) + x.amax(dim=(1, 2)) * torch.from_numpy(op.w_max))
                        * torch.from_numpy(op.scale) + torch.from_numpy(op.shift))                              # [n, c]
        m = torch.stack([x.mean(dim=-1), x.amax(dim=-1)], dim=1)
        sp = torch.relu(F.conv2d(m, torch.from_numpy(op.w_sp)[None], torch.tensor([op.b_sp]), padding=2))[:, 0]   # [n, h, w]
        T[op.dst][..., : op.c] = _bf16(torch.sigmoid(ch[:, None, None, :] * sp[..., None]) * x)
        return None
    if isinstance(op, L.CsamGateSpec):
        import torch.nn.functional as F
        x = T[op.src][..., : op.c]
        z = (x.mean(dim=(1, 2)) * torch.from_numpy(op.w_avg) + x.amax(dim=(1, 2)) * torch.from_numpy(op.w_max)) * torch.from_numpy(op.scale) \
            + torch.from_numpy(op.shift)
        ch = torch.relu(F.gelu(z) @ torch.from_numpy(op.l1).T) @ torch.from_numpy(op.l2).T + torch.from_numpy(op.b2)        # [n, c]
        m = torch.stack([x.mean(dim=-1), x.amax(dim=-1)], dim=1)
        sp = F.conv2d(torch.relu(F.conv2d(m, torch.from_numpy(op.w21)[None], None, padding=1)), torch.from_numpy(op.w22)[None, None], None,
                      padding=1)[:, 0]
        g = torch.sigmoid(ch)[:, None, None, :] + torch.sigmoid(sp)[..., None]
        T[op.dst][..., : op.c] = _bf16(g * x * torch.from_numpy(op.bt_scale) + torch.from_numpy(op.bt_shift))
        return None
    if isinstance(op, L.VffmSpec):
        low, high, mixed, local = T[op.low], T[op.high], T[op.mixed], T[op.local]
        g = 0.0
        for v, b in zip((mixed.mean(dim=(1, 2)), mixed.amax(dim=(1, 2))), op.branches):
            hid = torch.relu(v @ torch.from_numpy(b["w1"]).T * torch.from_numpy(b["s1"]) + torch.from_numpy(b["t1"]))
            g = g + (hid @ torch.from_numpy(b["w2"]).T * torch.from_numpy(b["s2"]) + torch.from_numpy(b["t2"]))
        wei = torch.sigmoid(g[:, None, None, :] + local)
        T[op.dst][...] = _bf16(2 * low * wei + 2 * high * (1 - wei))
        return None
    if isinstance(op, L.SumSpec):
        T[op.dst][...] = _bf16(sum(T[s_] for s_ in op.srcs))
        return None
    if isinstance(op, L.ChannelGateSpec):
        x = T[op.src][..., : op.c]
        g = torch.sigmoid(torch.relu(x.mean(dim=(1, 2)) @ torch.from_numpy(op.w1).T) @ torch.from_numpy(op.w2).T)[:, None, None, :]
        if op.mode == 0:
            y = x * g + (T[op.res][..., : op.c] if op.res is not None else 0.0)
            y = torch.relu(y)
        else:
            gs = torch.sigmoid((x * torch.from_numpy(op.ws)).sum(-1, keepdim=True))
            y = x * (1.0 + g + gs)
        y = _bf16(y)
        T[op.dst][..., : op.c] = y
        if op.dst_s2d is not None:
            for py in range(2):
                for px in range(2):
                    k = (py * 2 + px) * op.c
                    T[op.dst_s2d][..., k: k + op.c] = y[:, py::2, px::2]
        return None
    if isinstance(op, L.LayerNormSpec):
        return run_layernorm(op, T)
    if isinstance(op, L.AttentionSpec):
        return run_attention(op, T)
    if isinstance(op, L.DWConvSpec):
        return run_dwconv(op, T)
    if isinstance(op, L.GraphConvSpec):
        return run_graph_conv(op, T)
    if isinstance(op, L.BilinearUpSpec):
        return run_bilinear_up(op, T)
    if isinstance(op, L.EcamHeadSpec):
        return run_ecam_head(op, T, chunk, ext, nv)
    if isinstance(op, L.MaxPoolS2DSpec):
        return run_maxpool_s2d(op, T)
    if isinstance(op, L.SegHeadSpec):
        return run_seg_head(op, T, chunk, ext, nv)
    raise TypeError(f"emulator: unknown op {op!r}")


def run_bit_transformer(op: L.BitTransformerSpec, T: Dict[str, torch.Tensor], chunk: int) -> None:
    """models/networks.py:359-394,414-428 on the bf16 feature map, fp32 arithmetic, one rounding at the output."""
    import torch.nn.functional as F
    x = T[op.src][..., : op.c]                                   # [2*chunk, h, w, c]
    n2, h, w, c = x.shape
    xf = x.reshape(n2, h * w, c)
    a = torch.softmax(xf @ torch.from_numpy(op.conv_a).T, dim=1)  # softmax over pixels, [2*chunk, hw, L]
    tok = torch.einsum("bnl,bnc->blc", a, xf)                    # [2*chunk, L, c]
    t = torch.cat([tok[:chunk], tok[chunk:]], dim=1) + torch.from_numpy(op.pos)[None]      # [chunk, 2L, c]
    scale = c ** -0.5

    def attend(q, k, v, softmax=True):
        b, n, inner = q.shape
        d = inner // op.heads
        qh, kh, vh = (z.reshape(b, -1, op.heads, d).transpose(1, 2) for z in (q, k, v))
        dots = torch.einsum("bhid,bhjd->bhij", qh, kh) * scale
        at = dots.softmax(-1) if softmax else dots
        return torch.einsum("bhij,bhjd->bhid", at, vh).transpose(1, 2).reshape(b, n, inner)

    for row in op.enc:
        P = L.bit_unpack(L.bit_enc_fields(c, op.inner_enc, op.mlp), row)
        y = F.layer_norm(t, (c,), P["ln1_g"], P["ln1_b"])
        q, k, v = (y @ P["wqkv"].T).chunk(3, dim=-1)
        t = attend(q, k, v) @ P["wout"].T + P["bout"] + t
        y = F.layer_norm(t, (c,), P["ln2_g"], P["ln2_b"])
        t = F.gelu(y @ P["w1"].T + P["b1"]) @ P["w2"].T + P["b2"] + t
    m = torch.cat([t[:, : op.token_len], t[:, op.token_len:]], dim=0)                     # [2*chunk, L, c]: per image
    z = xf
    for row in op.dec:
        P = L.bit_unpack(L.bit_dec_fields(c, op.inner_dec, op.mlp), row)
        zn, mn = F.layer_norm(z, (c,), P["ln1_g"], P["ln1_b"]), F.layer_norm(m, (c,), P["ln1_g"], P["ln1_b"])
        o = attend(zn @ P["wq"].T, mn @ P["wk"].T, mn @ P["wv"].T, op.softmax)
        z = o @ P["woutt"] + P["bout"] + z
        y = F.layer_norm(z, (c,), P["ln2_g"], P["ln2_b"])
        z = F.gelu(y @ P["w1t"] + P["b1"]) @ P["w2t"] + P["b2"] + z
    T[op.dst][..., : c] = _bf16(z.reshape(n2, h, w, c))
