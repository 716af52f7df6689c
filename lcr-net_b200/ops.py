"""Host-side operator layer: the reference's ``experiments/lcrnet/modules/ops`` wrappers
(grid_subsample.py:7-22, radius_search.py:7-27) plus thin torch-tensor wrappers over the C ABI
for the model kernels.  All tensors are CUDA tensors; nothing here computes on the CPU."""
import torch

from . import _lib, ext

GROUPS = 32


def use_tensor_cores():
    """GEMMs run on the tcgen05 3xTF32 kernel (gemm_tc.cu) unless LCR_GEMM=simt selects the fp32 SIMT
    kernel (gemm.cu); both are fp32-accurate."""
    import os
    return os.environ.get('LCR_GEMM', 'tc') != 'simt'


def grid_subsample(points, lengths, voxel_size, order='reference'):
    """ops/grid_subsample.py:7-22."""
    s_points, s_lengths = ext.grid_subsampling(points, lengths, voxel_size, order=order)
    return s_points, s_lengths


def radius_search(q_points, s_points, q_lengths, s_lengths, radius, neighbor_limit, int32=False, defer=None, grid=None,
                  nearest=False):
    """ops/radius_search.py:7-27 (the ``[:, :neighbor_limit]`` cut is fused into the kernel).
    ``defer`` / ``grid`` / ``nearest``: see ext.radius_neighbors."""
    return ext.radius_neighbors(q_points, s_points, q_lengths, s_lengths, radius,
                                limit=neighbor_limit if neighbor_limit and neighbor_limit > 0 else 0, int32=int32,
                                defer=defer, grid=grid, nearest=nearest)


def _f32c(t):
    assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous(), 'expected a contiguous CUDA float tensor'
    return t


class Derived:
    """A tensor (or tuple of tensors) derived from model weights by kernels queued on the stream that was
    current when it was built.  The pipelines run chunks on several CUDA streams from several host threads
    and share these cached values: ``get()`` makes the consuming stream wait for the producing kernels
    (a recorded event), so a second stream can never read halves that are still only queued."""

    __slots__ = ('key', 'value', 'stream', 'event')

    def __init__(self, key, value):
        self.key, self.value = key, value
        if torch.cuda.is_available():
            self.stream = torch.cuda.current_stream()
            self.event = torch.cuda.Event()
            self.event.record(self.stream)
        else:
            self.stream = self.event = None

    def get(self):
        if self.event is not None:
            cur = torch.cuda.current_stream()
            if cur != self.stream:
                cur.wait_event(self.event)
        return self.value


def derived(owner, slot, key, build):
    """Cached ``build()`` on ``owner.<slot>``, rebuilt when ``key`` changes (weights reloaded / updated in place)."""
    ent = getattr(owner, slot, None)
    if ent is None or ent.key != key:
        ent = Derived(key, build())
        setattr(owner, slot, ent)
    return ent.get()


_SPLIT_CACHE = {}


def tf32_split(weight):
    """(hi, lo) tf32 halves of a weight tensor for the 3xTF32 GEMM (lcr_tf32_split), cached per
    (storage, version): the cache entry keeps ``weight`` alive, so its address cannot be reused by
    another tensor while the entry exists; an in-place update bumps the version and re-splits.
    The entry carries the event of its split kernel (see Derived): safe to share between streams."""
    key = (weight.data_ptr(), weight._version, tuple(weight.shape), tuple(weight.stride()))
    ent = _SPLIT_CACHE.get(key)
    if ent is None:
        if len(_SPLIT_CACHE) > 4096:
            _SPLIT_CACHE.clear()
        w = weight.detach()
        assert w.is_cuda and w.dtype == torch.float32 and w.stride(-1) == 1
        wc = w.contiguous()
        halves = torch.empty((2,) + tuple(wc.shape), dtype=torch.float32, device=w.device)
        _lib.check(_lib.lib().lcr_tf32_split(_lib.ptr(wc), wc.numel(), _lib.ptr(halves[0]), _lib.ptr(halves[1]),
                                             _lib.stream_ptr(w.device)))
        ent = Derived(key, (weight, halves[0], halves[1]))
        _SPLIT_CACHE[key] = ent
    v = ent.get()
    return v[1], v[2]


def prepare(net, sync=True):
    """Builds every derived weight (transposes, packed BatchNorm parameters, fused QKV weights, tf32 halves)
    of ``net`` once, on the current stream, then synchronises: called by the multi-stream pipelines before
    the first chunk is queued so no stream ever builds -- or half-builds -- a cache entry another one reads."""
    for m in net.modules():
        fn = getattr(m, 'prepare_b200', None)
        if fn is not None:
            fn()
    if sync and torch.cuda.is_available():
        torch.cuda.current_stream().synchronize()
    return net


def _host_ptr(t):
    if t is None:
        return None
    assert not t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.numel() == 45
    return _lib.ptr(t)


def as_index32(idx):
    """Neighbour tables are int32 inside the kernels; the reference's int64 tables are narrowed."""
    if idx.dtype != torch.int32:
        idx = idx.to(torch.int32)
    if idx.stride(-1) != 1:
        idx = idx.contiguous()
    return idx


def row_flags(x):
    _lib.require_cuda(x)
    flags = torch.empty(x.shape[0], dtype=torch.uint8, device=x.device)
    _lib.check(_lib.lib().lcr_row_flags(_lib.ptr(_f32c(x)), x.shape[0], x.shape[1], _lib.ptr(flags),
                                        _lib.stream_ptr(x.device)))
    return flags


def gn_fusable(stacks):
    """GroupNorm statistics can ride in the tensor-core GEMM epilogue (gemm_tc.cu GnFuse) when every stack
    has at least one 32-row block to itself; LCR_GN_FUSE=0 selects the separate statistics pass."""
    import os
    return (stacks is not None and stacks.min_rows >= 32 and use_tensor_cores()
            and os.environ.get('LCR_GN_FUSE', '1') != '0')


def _finalize_blocks(partial, rows, channels, stacks, eps, groups, device):
    stats = torch.empty((stacks.n, groups, 2), dtype=torch.float32, device=device)
    _lib.check(_lib.lib().lcr_group_norm_finalize_blocks(_lib.ptr(partial), rows, channels, groups,
                                                         _lib.ptr(stacks.off), stacks.n, eps, _lib.ptr(stats),
                                                         _lib.stream_ptr(device)))
    return stats


def kpconv(s_feats, q_points, s_points, idx, kernel_points, sigma, weights, bias, s_flags=None, weights_nk=None,
           kernel_points_host=None, gn=None):
    """KPConv.forward (kpconv.py:79-122).  ``s_flags``: row-sum>0 flags of s_feats (computed if None).
    ``kernel_points_host``: contiguous float32 CPU copy of kernel_points (selects the sparse gather).
    ``gn`` = (stacks, eps, groups): also return the GroupNorm statistics of the output -> (out, stats)."""
    _lib.require_cuda(s_feats, q_points, s_points, idx)
    L = _lib.lib()
    idx = as_index32(idx)
    m, n = q_points.shape[0], s_points.shape[0]
    c_in, c_out = weights.shape[1], weights.shape[2]
    if s_flags is None and c_in > 1:
        s_flags = row_flags(s_feats)
    out = torch.empty((m, c_out), dtype=torch.float32, device=s_feats.device)
    w_hi = w_lo = None
    if weights_nk is not None and c_in > 1 and use_tensor_cores():
        w_hi, w_lo = tf32_split(weights_nk)
    ws_bytes = L.lcr_kpconv_ws_bytes2(m, n, c_in)
    ws = _lib.workspace.get(ws_bytes, s_feats.device, slot=1)
    if gn is not None and w_hi is not None and gn_fusable(gn[0]) and m > 0:
        stacks, eps, groups = gn
        assert stacks.rows == m
        part = _lib.workspace.get(L.lcr_gn_blocks_ws_bytes(m, c_out), s_feats.device, slot=4)
        _lib.check(L.lcr_kpconv_gn(_lib.ptr(_f32c(s_feats)), _lib.ptr(s_flags), n, _lib.ptr(_f32c(q_points)), m,
                                   _lib.ptr(_f32c(s_points)), _lib.ptr(idx), idx.stride(0), idx.shape[1],
                                   _lib.ptr(_f32c(kernel_points)), _host_ptr(kernel_points_host), float(sigma),
                                   _lib.ptr(_f32c(weights)), _lib.ptr(w_hi), _lib.ptr(w_lo), _lib.ptr(bias), c_in, c_out,
                                   _lib.ptr(out), _lib.ptr(ws), ws.numel(), _lib.ptr(part), _lib.ptr(stacks.off),
                                   stacks.n, _lib.stream_ptr(s_feats.device)))
        return out, _finalize_blocks(part, m, c_out, stacks, eps, groups, s_feats.device)
    _lib.check(L.lcr_kpconv(_lib.ptr(_f32c(s_feats)), _lib.ptr(s_flags), n, _lib.ptr(_f32c(q_points)), m,
                            _lib.ptr(_f32c(s_points)), _lib.ptr(idx), idx.stride(0), idx.shape[1],
                            _lib.ptr(_f32c(kernel_points)), _host_ptr(kernel_points_host), float(sigma), _lib.ptr(_f32c(weights)),
                            _lib.ptr(w_hi), _lib.ptr(w_lo), _lib.ptr(bias), c_in, c_out, _lib.ptr(out), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(s_feats.device)))
    if gn is not None:
        return out, group_norm_stats(out, gn[0], gn[1], gn[2])
    return out


def linear(x, weight_t, bias, weight_nk=None, gn=None):
    """nn.Linear: weight_t = weight transposed to [c_in, c_out] (SIMT kernel); weight_nk = the
    nn.Linear layout [c_out, c_in] (tensor-core kernel, used when c_in % 32 == 0).
    ``gn`` = (stacks, eps, groups): also return the GroupNorm statistics of the output -> (out, stats)."""
    _lib.require_cuda(x)
    if gn is not None:
        stacks, eps, groups = gn
        if (weight_nk is not None and x.shape[1] % 32 == 0 and weight_nk.shape[0] % 4 == 0 and gn_fusable(stacks)
                and x.shape[0] > 0):
            assert stacks.rows == x.shape[0]
            L = _lib.lib()
            c_out = weight_nk.shape[0]
            out = torch.empty((x.shape[0], c_out), dtype=torch.float32, device=x.device)
            w_hi, w_lo = tf32_split(weight_nk)
            part = _lib.workspace.get(L.lcr_gn_blocks_ws_bytes(x.shape[0], c_out), x.device, slot=4)
            _lib.check(L.lcr_linear_tc_gn(_lib.ptr(_f32c(x)), x.shape[0], x.shape[1], x.stride(0), _lib.ptr(w_hi),
                                          _lib.ptr(w_lo), c_out, w_hi.stride(0), _lib.ptr(bias), _lib.ptr(out),
                                          out.stride(0), _lib.ptr(part), _lib.ptr(stacks.off), stacks.n,
                                          _lib.stream_ptr(x.device)))
            return out, _finalize_blocks(part, x.shape[0], c_out, stacks, eps, groups, x.device)
        out = linear(x, weight_t, bias, weight_nk)
        return out, group_norm_stats(out, stacks, eps, groups)
    if weight_nk is not None and x.shape[1] % 32 == 0 and weight_nk.shape[0] % 4 == 0 and use_tensor_cores():
        out = torch.empty((x.shape[0], weight_nk.shape[0]), dtype=torch.float32, device=x.device)
        w_hi, w_lo = tf32_split(weight_nk)
        _lib.check(_lib.lib().lcr_linear_tc(_lib.ptr(_f32c(x)), x.shape[0], x.shape[1], x.stride(0),
                                            _lib.ptr(w_hi), _lib.ptr(w_lo), weight_nk.shape[0], w_hi.stride(0),
                                            _lib.ptr(bias), None, 0, _lib.ptr(out), out.stride(0),
                                            _lib.stream_ptr(x.device)))
        return out
    out = torch.empty((x.shape[0], weight_t.shape[1]), dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib().lcr_linear(_lib.ptr(_f32c(x)), x.shape[0], x.shape[1], _lib.ptr(_f32c(weight_t)),
                                     weight_t.shape[1], _lib.ptr(bias), _lib.ptr(out), _lib.stream_ptr(x.device)))
    return out


def host_to_device(t, device):
    """Small host tensor -> device without stalling the host.  ``torch.tensor(..., device='cuda')`` (and ``.to(dev)``)
    copy synchronously: torch waits for every kernel queued on the stream -- five times per descriptor forward,
    0.9 ms each.  A ``non_blocking`` copy from pageable memory returns as soon as the driver has staged the bytes (no
    device synchronisation, the source may be freed at once), and needs no pinned allocation."""
    if torch.device(device).type != 'cuda':
        return t
    return t.to(device, non_blocking=True)


class Stacks:
    """Row offsets of the stacks (units of one reference forward) at one pyramid level."""

    def __init__(self, lengths_host, device):
        import itertools
        off = [0] + list(itertools.accumulate(int(x) for x in lengths_host))
        self.n = len(off) - 1
        self.rows = off[-1]
        self.max_rows = max([b - a for a, b in zip(off[:-1], off[1:])] + [1])
        self.min_rows = min([b - a for a, b in zip(off[:-1], off[1:])] + [self.rows])
        self.off_host = off
        self.off = host_to_device(torch.tensor(off, dtype=torch.int64), device)


def group_norm_stats(x, stacks, eps=1e-5, groups=GROUPS):
    L = _lib.lib()
    stats = torch.empty((stacks.n, groups, 2), dtype=torch.float32, device=x.device)
    ws_bytes = L.lcr_group_norm_ws_bytes(stacks.max_rows, stacks.n, groups)
    ws = _lib.workspace.get(ws_bytes, x.device, slot=2)
    _lib.check(L.lcr_group_norm_stats(_lib.ptr(_f32c(x)), x.shape[0], x.shape[1], groups, _lib.ptr(stacks.off),
                                      stacks.n, stacks.max_rows, eps, _lib.ptr(stats), _lib.ptr(ws), ws.numel(),
                                      _lib.stream_ptr(x.device)))
    return stats


def group_norm_apply(x, stats, gamma, beta, stacks, leaky=True, other=None, other_norm=None, want_flags=False,
                     groups=GROUPS, slope=0.1):
    """y = act(gn(x) + other); other_norm = (stats2, gamma2, beta2) normalises ``other`` first."""
    y = torch.empty_like(x)
    flags = torch.empty(x.shape[0], dtype=torch.uint8, device=x.device) if want_flags else None
    s2, g2, b2 = other_norm if other_norm is not None else (None, None, None)
    _lib.check(_lib.lib().lcr_group_norm_apply(
        _lib.ptr(_f32c(x)), _lib.ptr(stats), _lib.ptr(gamma), _lib.ptr(beta), _lib.ptr(other), _lib.ptr(s2),
        _lib.ptr(g2), _lib.ptr(b2), x.shape[0], x.shape[1], groups, _lib.ptr(stacks.off), stacks.n,
        1 if leaky else 0, slope, _lib.ptr(y), _lib.ptr(flags), _lib.stream_ptr(x.device)))
    return (y, flags) if want_flags else y


def maxpool(x, idx):
    """functional.py:54-67."""
    idx = as_index32(idx)
    out = torch.empty((idx.shape[0], x.shape[1]), dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib().lcr_maxpool(_lib.ptr(_f32c(x)), x.shape[0], _lib.ptr(idx), idx.stride(0), idx.shape[1],
                                      idx.shape[0], x.shape[1], _lib.ptr(out), _lib.stream_ptr(x.device)))
    return out


def netvlad(feats, scan_off, n_scans, cluster_weights, cluster_weights2, hidden1_weights, bn1, bn2, gating_weights,
            gating_bn):
    L = _lib.lib()
    out = torch.empty((n_scans, 256), dtype=torch.float32, device=feats.device)
    ws_bytes = L.lcr_netvlad_ws_bytes(feats.shape[0], n_scans)
    ws = _lib.workspace.get(ws_bytes, feats.device, slot=3)
    _lib.check(L.lcr_netvlad(_lib.ptr(_f32c(feats)), feats.shape[0], _lib.ptr(scan_off), n_scans,
                             _lib.ptr(_f32c(cluster_weights)), _lib.ptr(_f32c(cluster_weights2)),
                             _lib.ptr(_f32c(hidden1_weights)), _lib.ptr(bn1), _lib.ptr(bn2),
                             _lib.ptr(_f32c(gating_weights)), _lib.ptr(gating_bn), _lib.ptr(out), _lib.ptr(ws),
                             ws.numel(), _lib.stream_ptr(feats.device)))
    return out


def l2_topk(queries, db, k, valid_counts=None):
    """Exact squared-L2 top-k (eval_loop_detection_overlap_dataset.py:183-214).  Returns
    (d2 [nq,k] f32, idx [nq,k] i64), ascending; missing entries (+inf, -1)."""
    _lib.require_cuda(queries, db)
    nq = queries.shape[0]
    d2 = torch.empty((nq, k), dtype=torch.float32, device=queries.device)
    idx = torch.empty((nq, k), dtype=torch.int64, device=queries.device)
    if valid_counts is not None:
        valid_counts = valid_counts.to(device=queries.device, dtype=torch.int32).contiguous()
    _lib.check(_lib.lib().lcr_l2_topk(_lib.ptr(_f32c(queries)), nq, _lib.ptr(_f32c(db)), db.shape[0],
                                      queries.shape[1], k, _lib.ptr(valid_counts), _lib.ptr(d2), _lib.ptr(idx),
                                      _lib.stream_ptr(queries.device)))
    return d2, idx
