"""Descriptor database build and loop-candidate retrieval.

Reference flow: ``test_loop_detection.py:60-69`` writes one ``{seq}_{idx}.npz`` (key ``anc_global``)
per scan, ``eval_loop_detection_overlap_dataset.py:162-214`` reloads them and, for every query
i in [101, N-2], searches the exact squared-L2 top-50 among rows [0, i-100) with a freshly built
faiss index.  Here the descriptors stay in HBM, the database is sharded over ranks in contiguous
blocks (rank r owns scans [r*N/G, (r+1)*N/G)), one all-gather assembles it on every rank, and a
single brute-force top-k kernel answers all queries of the rank.

Multi-GPU: one process per GPU, ``torch.distributed`` (NCCL on GPUs; the host-side sharding
logic is backend-agnostic and is tested with gloo).
"""
import numpy as np
import torch
import torch.distributed as dist

from . import ops


def shard_range(n_total, rank, world):
    """Contiguous block of rank: global index = start + local index."""
    per = (n_total + world - 1) // world
    start = min(rank * per, n_total)
    return start, min(start + per, n_total)


def all_gather_descriptors(local, n_total=None, group=None):
    """local [n_local, 256] of this rank -> [n_total, 256] on every rank, in global scan order.
    A single all-gather when every rank holds the same number of rows (the sharding above pads
    the last shard), otherwise per-rank broadcasts."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    counts = [torch.zeros(1, dtype=torch.int64, device=local.device) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device), group=group)
    counts = [int(c) for c in counts]
    per = max(counts)
    if all(c == per for c in counts):
        out = torch.empty((world * per, local.shape[1]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    padded = torch.zeros((per, local.shape[1]), dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    out = torch.empty((world * per, local.shape[1]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    return torch.cat([out[r * per:r * per + c] for r, c in enumerate(counts)], 0)


def causal_valid_counts(query_ids, gap=100):
    """Query i searches database rows [0, i - gap) (eval_loop_detection_overlap_dataset.py:186-197:
    the index is filled with ``emb[:i-100]``)."""
    return np.maximum(np.asarray(query_ids, dtype=np.int64) - gap, 0).astype(np.int32)


def search(queries, db, k=25, valid_counts=None):
    """Exact squared-L2 top-k on the GPU (ops.l2_topk)."""
    return ops.l2_topk(queries, db, k, valid_counts)


def loop_candidates(db, query_ids=None, k=50, gap=100):
    """The reference evaluation's candidate rows (i, j, d2) with i - j >= gap
    (eval_loop_detection_overlap_dataset.py:183-214), for all queries at once."""
    n = db.shape[0]
    if query_ids is None:
        query_ids = np.arange(gap + 1, max(n - 1, gap + 1))
    query_ids = np.asarray(query_ids, dtype=np.int64)
    if len(query_ids) == 0:
        return np.zeros((0, 3), dtype=np.float64)
    q = db[torch.as_tensor(query_ids, device=db.device)]
    d2, idx = search(q, db, k, torch.from_numpy(causal_valid_counts(query_ids, gap)))
    d2, idx = d2.cpu().numpy(), idx.cpu().numpy()
    rows = [(int(i), int(j), float(d)) for i, dr, jr in zip(query_ids, d2, idx) for d, j in zip(dr, jr) if j >= 0]
    return np.array(rows, dtype=np.float64).reshape(-1, 3)


def build_database(model, scans, collate, batch_scans=32, rank=0, world=1, group=None):
    """Descriptors of this rank's shard of ``scans`` (a sequence of float32 [N,3] arrays, indexed
    globally) followed by the all-gather.  ``collate(list_of_scans) -> data_dict`` (see
    data.scans_collate_fn_stack_mode).  Returns (db [n_total,256] on every rank, (start, end))."""
    start, end = shard_range(len(scans), rank, world)
    out = []
    for b0 in range(start, end, batch_scans):
        batch = [scans[i] for i in range(b0, min(b0 + batch_scans, end))]
        out.append(model(collate(batch))['anc_global'])
    dev = next(model.parameters()).device
    local = torch.cat(out, 0) if out else torch.zeros((0, 256), dtype=torch.float32, device=dev)
    return all_gather_descriptors(local, len(scans), group), (start, end)


def save_descriptor_npz(path, descriptor):
    """The reference's on-disk descriptor record (test_loop_detection.py:60-69)."""
    np.savez_compressed(path, anc_global=np.asarray(descriptor, dtype=np.float32).reshape(1, 256))
