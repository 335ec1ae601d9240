"""K1 (csrc/conv_gemm.cuh) against the CPU emulator, op by op, through the C-ABI.

Each case is a one-op Program: source tensors are written with stcd_plan_tensor_copy, the op
runs via stcd_forward, outputs are read back and compared with oracle/emulate.py, which
executes the same K-program with fp32 accumulation.  Outputs are bf16, so the bar is one bf16
ulp (accumulation-order differences can flip a rounding); fp32 external outputs: 1e-3 relative.
"""
import numpy as np
import pytest
import torch

from oracle import emulate
from stcd_b200 import lowering as L

pytestmark = pytest.mark.gpu


def _rand_bf16(g, *shape):
    return torch.randn(*shape, generator=g).to(torch.bfloat16).to(torch.float32)


def _check_bf16(got, want, name):
    diff = (got - want).abs()
    tol = 0.0079 * want.abs() + 2e-3
    bad = (diff > tol).float().mean().item()
    assert bad == 0.0, f"{name}: {bad:.4%} elements beyond 1 bf16 ulp, max diff {diff.max().item():.4g}"


def _run_case(prog, chunk, seed=0, n_valid=None):
    from stcd_b200.plan import Plan
    g = torch.Generator().manual_seed(seed)
    plan = Plan(prog, chunk)
    outs_names = set()
    for op in prog.ops:
        for n in (op.out0, op.out_raw, op.out_pool, op.out_diff):
            if n is not None:
                outs_names.add(n)
    T = {}
    for name, t in prog.tensors.items():
        if name in outs_names:
            T[name] = torch.zeros(t.mult * chunk, t.h, t.w, t.c)
        else:
            T[name] = _rand_bf16(g, t.mult * chunk, t.h, t.w, t.c)
            plan.write_tensor(name, T[name])
    nv = chunk if n_valid is None else n_valid
    ext = [torch.zeros(chunk, e.channels, e.h, e.w) for e in prog.ext]
    for op in prog.ops:
        emulate.run_conv(op, T, chunk, ext, nv)
    got_ext = plan.run_raw(nv)
    for name in outs_names:
        _check_bf16(plan.read_tensor(name), T[name], name)
    for k, e in enumerate(prog.ext):
        a, b = got_ext[k].cpu(), ext[k][:nv]
        assert (a - b).abs().max().item() <= 1e-3 * (b.abs().max().item() + 1.0), f"ext {k}"
    plan.close()


def _w(g, cout, cin, k):
    return torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5


def _affine(g, cout):
    return (0.5 + torch.rand(cout, generator=g)).numpy(), (0.1 * torch.randn(cout, generator=g)).numpy()


def _prog(h, w):
    return L.Program(model="unit", in_channels=0, h=h, w=w)


@pytest.mark.parametrize("cin,cout,h,w,pair", [
    (16, 16, 24, 40, False),      # kc=16, ragged tiles in both directions
    (32, 32, 16, 32, True),       # kc=32, Siamese pair tiles
    (64, 128, 16, 16, False),     # kc=64, N=128
    (128, 256, 8, 16, False),     # N=256: the whole TMEM row of one accumulator
    (64, 320, 8, 16, False),      # cout > n_tile: several N tiles (blockIdx.y), padded cout
    (256, 64, 8, 16, True),       # long K-program
])
def test_conv3x3(cin, cout, h, w, pair):
    g = torch.Generator().manual_seed(cin * 1000 + cout)
    p = _prog(h, w)
    p.tensor("a", 2, h, w, cin)
    p.tensor("y", 2, h, w, cout)
    s, b = _affine(g, cout)
    L.add_conv(p, "c", [L.Segment("a", cin)], L.conv_taps(_w(g, cout, cin, 3), 1), cout, h, w,
               1 if pair else 2, s, b, pair=pair, relu=True, out0="y")
    _run_case(p, chunk=2)


def test_pair_epilogue_pool_and_absdiff():
    g = torch.Generator().manual_seed(1)
    h, w, c = 16, 32, 32
    p = _prog(h, w)
    p.tensor("a", 2, h, w, c)
    p.tensor("y", 2, h, w, c)
    p.tensor("pool", 2, h // 2, w // 2, c)
    p.tensor("d", 1, h, w, c)
    s, b = _affine(g, c)
    L.add_conv(p, "c", [L.Segment("a", c)], L.conv_taps(_w(g, c, c, 3), 1), c, h, w, 1, s, b, pair=True, relu=True,
               out0="y", out_pool="pool", out_diff="d")
    _run_case(p, chunk=3)


def test_nested_block_epilogue_raw_bn_residual():
    """SNUNet conv_block_nested (SNUNet.py:17-26): conv1 emits the pre-BN identity and relu(bn1(.));
    conv2 adds the identity before the ReLU."""
    g = torch.Generator().manual_seed(2)
    h, w, cin, c = 16, 16, 48, 32
    p = _prog(h, w)
    p.tensor("a", 1, h, w, cin)
    p.tensor("ident", 1, h, w, c)
    p.tensor("t", 1, h, w, c)
    p.tensor("y", 1, h, w, c)
    bias = (0.1 * torch.randn(c, generator=g)).numpy()
    s1, b1 = _affine(g, c)
    L.add_conv(p, "conv1", [L.Segment("a", cin)], L.conv_taps(_w(g, c, cin, 3), 1), c, h, w, 1,
               np.ones(c, np.float32), bias, scale2=s1, shift2=b1, relu=True, out_raw="ident", out0="t")
    s2, b2 = _affine(g, c)
    L.add_conv(p, "conv2", [L.Segment("t", c)], L.conv_taps(_w(g, c, c, 3), 1), c, h, w, 1, s2, b2, relu=True,
               res="ident", out0="y")
    # run the two ops one after the other: 't' and 'ident' are produced by op 1
    from stcd_b200.plan import Plan
    plan = Plan(p, 2)
    a = _rand_bf16(g, 2, h, w, cin)
    plan.write_tensor("a", a)
    plan.run_raw()
    T = {"a": a, "ident": torch.zeros(2, h, w, c), "t": torch.zeros(2, h, w, c), "y": torch.zeros(2, h, w, c)}
    for op in p.ops:
        emulate.run_conv(op, T, 2, [], 2)
    for n in ("ident", "t", "y"):
        _check_bf16(plan.read_tensor(n), T[n], n)
    plan.close()


def test_virtual_concat_mixed_sources_and_streams():
    g = torch.Generator().manual_seed(3)
    h, w = 16, 16
    p = _prog(h, w)
    p.tensor("a", 1, h, w, 32)
    p.tensor("b", 2, h, w, 64)
    p.tensor("c", 1, h, w, 16)
    p.tensor("y", 1, h, w, 48)
    s, b = _affine(g, 48)
    segs = [L.Segment("a", 32), L.Segment("b", 64, stream=1), L.Segment("c", 16), L.Segment("b", 64, stream=0)]
    L.add_conv(p, "cat", segs, L.conv_taps(_w(g, 48, 176, 3), 1), 48, h, w, 1, s, b, relu=False, out0="y")
    _run_case(p, chunk=2)


@pytest.mark.parametrize("k,s,pad,cin,cout", [(3, 2, 1, 32, 32), (2, 2, 0, 64, 64), (4, 2, 1, 16, 32)])
def test_conv_transpose_phases(k, s, pad, cin, cout):
    g = torch.Generator().manual_seed(k)
    h, w = 8, 16
    p = _prog(h, w)
    p.tensor("a", 1, h, w, cin)
    p.tensor("y", 1, h * 2, w * 2, cout)
    wt = torch.randn(cin, cout, k, k, generator=g) / (cin * k * k / 4) ** 0.5
    bias = (0.1 * torch.randn(cout, generator=g)).numpy()
    L.add_conv(p, "up", [L.Segment("a", cin)], L.convT_phase_taps(wt, s, pad), cout, h, w, 1,
               np.ones(cout, np.float32), bias, osy=2, osx=2, out0="y")
    _run_case(p, chunk=2)


@pytest.mark.parametrize("k,pad,cin,cout", [(3, 1, 64, 128), (1, 0, 64, 128), (7, 3, 16, 64)])
def test_strided_conv(k, pad, cin, cout):
    g = torch.Generator().manual_seed(k + 10)
    h, w = 32, 32
    p = _prog(h, w)
    p.tensor("a", 1, h, w, cin)
    p.tensor("y", 1, h // 2, w // 2, cout)
    s, b = _affine(g, cout)
    L.add_conv(p, "s2", [L.Segment("a", cin, sy=2, sx=2)], L.conv_taps(_w(g, cout, cin, k), pad), cout, h // 2, w // 2,
               1, s, b, relu=True, out0="y")
    _run_case(p, chunk=2)


@pytest.mark.parametrize("cout", [1, 2])
def test_external_fp32_logits_and_ragged_chunk(cout):
    g = torch.Generator().manual_seed(4 + cout)
    h, w = 16, 32
    p = _prog(h, w)
    p.tensor("a", 1, h, w, 16)
    bias = (0.1 * torch.randn(cout, generator=g)).numpy()
    L.add_conv(p, "head", [L.Segment("a", 16)], L.conv_taps(_w(g, cout, 16, 3), 1), cout, h, w, 1,
               np.ones(cout, np.float32), bias, out_ext=0)
    p.ext.append(L.ExtOutput("logits", cout, h, w))
    _run_case(p, chunk=3, n_valid=2)
