"""ViG Grapher graph ops on the GPU (csrc/graph_kernels.cuh) against oracle/gcn.py at the ChangeGNN stage
shapes: neighbour tables equal except where two candidate distances differ by fp32
rounding noise, max-relative features bit-exact given the same table."""
import pytest
import torch

from oracle import gcn as ogcn
from stcd_b200 import gcn

pytestmark = pytest.mark.gpu

# (C, H, reduce ratio r, k, dilation): the four ChangeGNN stages at 256x256 input (N = H*H, M = N / r^2)
STAGES = [(80, 64, 4, 9, 1), (160, 32, 2, 9, 1), (400, 16, 1, 9, 2), (640, 8, 1, 9, 3), (24, 10, 1, 4, 2)]


def _case(c, h, r, seed, batch=2):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, c, h, h, generator=g)
    # smooth the map a little: neighbouring nodes are similar, as after a conv
    x = x + 0.5 * torch.nn.functional.avg_pool2d(x, 3, 1, 1)
    y = torch.nn.functional.avg_pool2d(x, r, r) if r > 1 else None
    n = h * h
    m = n // (r * r)
    rp = 0.05 * torch.randn(1, n, m, generator=g)
    return x.reshape(batch, c, n, 1), None if y is None else y.reshape(batch, c, m, 1), rp


@pytest.mark.parametrize("c,h,r,k,d", STAGES)
def test_knn_graph_matches_oracle(c, h, r, k, d):
    x, y, rp = _case(c, h, r, seed=c + h)
    want = ogcn.dense_dilated_knn_graph(x, y, k, d, rp)
    got = gcn.DenseDilatedKnnGraph(k, d)(x.cuda(), None if y is None else y.cuda(), rp.cuda()).cpu()
    assert got.shape == want.shape and got.dtype == torch.int64
    assert torch.equal(got[1], want[1])
    same = (got[0] == want[0])
    frac = same.float().mean().item()
    assert frac >= 0.999, frac
    if frac < 1.0:
        # every disagreement must be a numerical tie: the two candidates' distances differ by fp32 noise
        xn = torch.nn.functional.normalize(x[..., 0], dim=1).transpose(1, 2)
        yn = xn if y is None else torch.nn.functional.normalize(y[..., 0], dim=1).transpose(1, 2)
        dist = ogcn.pairwise_distance(xn.double(), yn.double()) + rp.double()
        b, n, t = torch.nonzero(~same, as_tuple=True)
        da = dist[b, n, got[0][b, n, t]]
        dw = dist[b, n, want[0][b, n, t]]
        assert (da - dw).abs().max().item() < 1e-5


@pytest.mark.parametrize("c,h,r,k,d", STAGES)
def test_max_relative_bit_exact(c, h, r, k, d):
    x, y, rp = _case(c, h, r, seed=7 * c + h)
    e = ogcn.dense_dilated_knn_graph(x, y, k, d, rp)
    want = ogcn.max_relative(x, e, y)
    got = gcn.max_relative(x.cuda(), e.cuda(), None if y is None else y.cuda()).cpu()
    assert torch.equal(got, want)
    z = gcn.max_relative(x.cuda(), e.cuda(), None if y is None else y.cuda(), interleave=True).cpu()
    assert torch.equal(z, ogcn.mr_features(x, e, y))


def test_no_relative_pos_and_self_graph():
    x, _, _ = _case(32, 12, 1, seed=3)
    got = gcn.DenseDilatedKnnGraph(5, 1)(x.cuda()).cpu()
    want = ogcn.dense_dilated_knn_graph(x, None, 5, 1)
    assert (got[0] == want[0]).float().mean().item() >= 0.999
    assert (got[0, :, :, 0] == torch.arange(144)).all(), "nearest node of a node is itself"


def test_argument_errors():
    x, y, rp = _case(16, 8, 2, seed=1)
    with pytest.raises(RuntimeError):
        gcn.DenseDilatedKnnGraph(3, 1)(x)                     # CPU tensor: no CPU path
    with pytest.raises(TypeError):
        gcn.DenseDilatedKnnGraph(3, 1)(x.cuda().half())
    from stcd_b200 import StcdError
    with pytest.raises(StcdError):
        gcn.DenseDilatedKnnGraph(9, 3)(x.cuda(), y.cuda())    # k*dilation = 27 > M = 16
    big = torch.randn(1, 8, 1024, 1).cuda()
    with pytest.raises(StcdError):
        gcn.DenseDilatedKnnGraph(3, 1)(big)                   # M = 1024 > 256 keys
