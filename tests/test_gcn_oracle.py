"""oracle/gcn.py (restatement of the absent gcn_lib; parity unpinned) against brute-force definitions."""
import torch

from oracle import gcn


def test_knn_graph_matches_bruteforce_definition():
    g = torch.Generator().manual_seed(11)
    x = torch.randn(2, 12, 36, 1, generator=g)
    y = torch.nn.functional.avg_pool2d(x.reshape(2, 12, 6, 6), 2, 2).reshape(2, 12, -1, 1)
    rp = 0.1 * torch.randn(1, 36, 9, generator=g)
    k, d = 2, 3
    e = gcn.dense_dilated_knn_graph(x, y, k, d, rp)
    assert e.shape == (2, 2, 36, k) and e.dtype == torch.int64
    xn = torch.nn.functional.normalize(x[..., 0], dim=1)
    yn = torch.nn.functional.normalize(y[..., 0], dim=1)
    for b in range(2):
        for i in range(36):
            dist = ((xn[b, :, i, None] - yn[b]) ** 2).sum(0) + rp[0, i]
            order = torch.argsort(dist)[: k * d: d]
            assert e[0, b, i].tolist() == order.tolist()
            assert e[1, b, i].tolist() == [i] * k


def test_max_relative_and_interleave():
    g = torch.Generator().manual_seed(12)
    x = torch.randn(2, 5, 16, 1, generator=g)
    e = gcn.dense_dilated_knn_graph(x, None, 3, 1)
    assert (e[0, :, :, 0] == torch.arange(16)).all(), "the nearest node of a node is itself when y is None"
    m = gcn.max_relative(x, e)
    for b in range(2):
        for n in range(16):
            want = (x[b, :, e[0, b, n], 0] - x[b, :, n, :]).max(dim=1).values
            assert torch.equal(m[b, :, n, 0], want)
    z = gcn.mr_features(x, e)
    assert torch.equal(z[:, 0::2], x) and torch.equal(z[:, 1::2], m)
