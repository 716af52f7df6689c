"""CPU, world_size 2, gloo: the host-side sharding / all-gather logic of the descriptor-database
build (the N>1 path).  The top-k itself is a CUDA kernel; here the gathered database is checked
with the numpy oracle."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_total, tmp):
    sys.path.insert(0, REPO)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from lcrnet_b200 import retrieval
    rng = np.random.default_rng(0)
    full = rng.standard_normal((n_total, 256)).astype(np.float32)
    s, e = retrieval.shard_range(n_total, rank, world)
    local = torch.from_numpy(full[s:e].copy())
    db = retrieval.all_gather_descriptors(local, n_total)
    ok = db.shape == (n_total, 256) and np.array_equal(db.numpy(), full)
    np.save(os.path.join(tmp, 'ok_%d.npy' % rank), np.array([ok, s, e]))
    dist.destroy_process_group()


@pytest.mark.parametrize('n_total', [64, 37])
def test_sharded_all_gather_gloo(tmp_path, n_total):
    port = 29500 + (os.getpid() + n_total) % 2000
    mp.spawn(_worker, args=(2, port, n_total, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = (np.load(tmp_path / ('ok_%d.npy' % r)) for r in (0, 1))
    assert r0[0] and r1[0]
    assert r0[1] == 0 and r0[2] == r1[1] and r1[2] == n_total      # contiguous cover


def test_shard_range_and_causal_counts():
    from lcrnet_b200 import retrieval
    cover = [retrieval.shard_range(32000, r, 8) for r in range(8)]
    assert cover[0] == (0, 4000) and cover[-1] == (28000, 32000)
    assert retrieval.shard_range(10, 3, 4) == (9, 10) and retrieval.shard_range(3, 3, 4) == (3, 3)
    assert list(retrieval.causal_valid_counts([0, 100, 101, 250])) == [0, 0, 1, 150]


def test_candidate_rows_semantics_with_oracle():
    """rows (i, j, d2) of the reference eval loop, via the numpy oracle."""
    from lcrnet_b200 import retrieval
    from oracle import model_oracle as mo
    rng = np.random.default_rng(1)
    db = rng.standard_normal((160, 256)).astype(np.float32)
    q_ids = np.arange(101, 159)
    d2, idx = mo.l2_topk(db[q_ids], db, 50, valid_counts=retrieval.causal_valid_counts(q_ids))
    assert ((q_ids[:, None] - idx)[idx >= 0] >= 100).all()
    assert (idx[0] >= 0).sum() == 1            # query 101 sees exactly one database row
