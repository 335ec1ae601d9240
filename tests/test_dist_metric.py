"""The N>1 path on the CPU (gloo, world size 2): pairs are sharded by rank, each rank accumulates its
own integer confusion matrix, and the path's only collective — the all-reduce of that matrix — must
reproduce the single-process matrix bit for bit, including uneven and empty shards."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import metric as ometric
from stcd_b200 import parallel


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _data(n_pairs, h=24, w=40, seed=11):
    g = np.random.default_rng(seed)
    pred = (g.random((n_pairs, h, w)) < 0.3).astype(np.uint8)
    label = (g.random((n_pairs, h, w)) < 0.1).astype(np.int64)
    return pred, label


def _worker(rank, world, port, n_pairs, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pred, label = _data(n_pairs)
        b, e = parallel.shard_range(n_pairs, rank, world)
        cm = torch.zeros(4, dtype=torch.int64)
        if e > b:
            cm += torch.from_numpy(ometric.confusion_matrix(pred[b:e], label[b:e]).reshape(-1).astype(np.int64))
        total = parallel.allreduce_confusion(cm, pixels=(e - b) * pred[0].size)
        assert int(total) == n_pairs * pred[0].size          # the pixel count rides in the same all-reduce
        assert int(cm.sum()) == int(total)
        assert parallel.rank_world() == (rank, world)
        np.save(os.path.join(out_dir, f"cm{rank}.npy"), cm.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_pairs", [8, 5, 1])
def test_sharded_confusion_matrix_equals_single_process(tmp_path, n_pairs):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_pairs, str(tmp_path)), nprocs=world, join=True)
    pred, label = _data(n_pairs)
    want = ometric.confusion_matrix(pred, label).reshape(-1).astype(np.int64)
    for r in range(world):
        got = np.load(tmp_path / f"cm{r}.npy")
        assert np.array_equal(got, want), (r, got, want)
    assert want.sum() == pred.size


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 8, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard_range(4, 2, 2)
    with pytest.raises(TypeError):
        parallel.allreduce_confusion(torch.zeros(4))
