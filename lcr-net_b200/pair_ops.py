"""torch-tensor wrappers over the C ABI for the registration-path kernels (pair.cu, matching.cu)."""
import torch

from . import _lib
from .ops import _f32c, as_index32


def _L():
    return _lib.lib()


def _s(t):
    return _lib.stream_ptr(t.device)


def linear_ex(x, weight_t, bias=None, relu=False, rowscale=None, out=None):
    """act(rowscale * (x . weight_t) + bias); x may be a column slice view (row stride = ld)."""
    assert x.stride(1) == 1
    n, cin = x.shape
    cout = weight_t.shape[1]
    if out is None:
        out = torch.empty((n, cout), dtype=torch.float32, device=x.device)
    _lib.check(_L().lcr_linear_ex(_lib.ptr(x), n, cin, x.stride(0), _lib.ptr(_f32c(weight_t)), cout, _lib.ptr(bias),
                                  _lib.ptr(rowscale), 1 if relu else 0, _lib.ptr(out), out.stride(0), _s(x)))
    return out


def linear_tc(x, weight, bias=None, relu=False, rowscale=None, out=None):
    """Tensor-core (tcgen05, 3xTF32) linear: weight in nn.Linear layout [c_out, c_in]."""
    assert x.stride(1) == 1 and weight.stride(1) == 1
    n, cin = x.shape
    cout = weight.shape[0]
    if out is None:
        out = torch.empty((n, cout), dtype=torch.float32, device=x.device)
    from .ops import tf32_split
    w_hi, w_lo = tf32_split(weight)
    _lib.check(_L().lcr_linear_tc(_lib.ptr(x), n, cin, x.stride(0), _lib.ptr(w_hi), _lib.ptr(w_lo), cout, w_hi.stride(0),
                                  _lib.ptr(bias), _lib.ptr(rowscale), 1 if relu else 0, _lib.ptr(out), out.stride(0),
                                  _s(x)))
    return out


def layer_norm(x, gamma, beta, residual=None, relu=False, eps=1e-5):
    y = torch.empty_like(x)
    _lib.check(_L().lcr_layer_norm(_lib.ptr(_f32c(x)), _lib.ptr(residual), _lib.ptr(gamma), _lib.ptr(beta), x.shape[0],
                                   x.shape[1], eps, 1 if relu else 0, _lib.ptr(y), _s(x)))
    return y


def rope_(x, theta):
    """In-place rotary embedding on a [rows, >=128] view (row stride = ld)."""
    assert x.stride(1) == 1 and theta.is_contiguous() and theta.shape[1] == 64
    _lib.check(_L().lcr_rope(_lib.ptr(x), x.stride(0), _lib.ptr(theta), x.shape[0], _s(x)))
    return x


def attention(q, k, v, q_off, k_off, n_problems, max_q_rows, heads=4, flops=0.0):
    """softmax(q k^T / sqrt(32)) v per head and problem.  q/k/v: [rows, heads*32] views."""
    out = torch.empty((q.shape[0], heads * 32), dtype=torch.float32, device=q.device)
    import os
    impl = os.environ.get('LCR_ATTN', 'tma')      # tma (default): tcgen05 + tensor-map TMA; tc: tcgen05 + per-row bulk
    if impl == 'tma':                              # copies; simt: fp32 flash-style kernel
        _lib.check(_L().lcr_attention_tma(_lib.ptr(q), q.stride(0), q.shape[0], _lib.ptr(k), k.stride(0), _lib.ptr(v),
                                          v.stride(0), k.shape[0], _lib.ptr(q_off), _lib.ptr(k_off), n_problems,
                                          max_q_rows, heads, 32, _lib.ptr(out), out.stride(0), float(flops), _s(q)))
        return out
    fn = _L().lcr_attention if impl == 'simt' else _L().lcr_attention_tc
    _lib.check(fn(_lib.ptr(q), q.stride(0), _lib.ptr(k), k.stride(0), _lib.ptr(v), v.stride(0),
                                  _lib.ptr(q_off), _lib.ptr(k_off), n_problems, max_q_rows, heads, 32, _lib.ptr(out),
                                  out.stride(0), float(flops), _s(q)))
    return out


def vote_shift(points, offsets, max_range):
    out = torch.empty_like(points)
    _lib.check(_L().lcr_vote_shift(_lib.ptr(_f32c(points)), _lib.ptr(offsets), offsets.stride(0), float(max_range),
                                   points.shape[0], _lib.ptr(out), _s(points)))
    return out


def nms_greedy(points, cloud_off, n_clouds, max_rows, radius):
    """-> keep u8 [N], counts i32 [n_clouds], kept_idx i32 [N] (compacted per cloud at its offset)."""
    n = points.shape[0]
    keep = torch.empty(n, dtype=torch.uint8, device=points.device)
    counts = torch.empty(n_clouds, dtype=torch.int32, device=points.device)
    kept_idx = torch.empty(max(n, 1), dtype=torch.int32, device=points.device)
    _lib.check(_L().lcr_nms_greedy(_lib.ptr(_f32c(points)), _lib.ptr(cloud_off), n_clouds, max_rows, float(radius),
                                   _lib.ptr(keep), _lib.ptr(counts), _lib.ptr(kept_idx), _s(points)))
    return keep, counts, kept_idx


def neighbor_mean(points, idx):
    idx = as_index32(idx)
    out = torch.empty((idx.shape[0], 3), dtype=torch.float32, device=points.device)
    _lib.check(_L().lcr_neighbor_mean(_lib.ptr(_f32c(points)), points.shape[0], _lib.ptr(idx), idx.stride(0),
                                      idx.shape[1], idx.shape[0], _lib.ptr(out), _s(points)))
    return out


def upsample_concat(coarse, up_idx, fine):
    up_idx = as_index32(up_idx)
    out = torch.empty((fine.shape[0], coarse.shape[1] + fine.shape[1]), dtype=torch.float32, device=fine.device)
    _lib.check(_L().lcr_upsample_concat(_lib.ptr(_f32c(coarse)), coarse.shape[0], coarse.shape[1], _lib.ptr(up_idx),
                                        up_idx.stride(0), _lib.ptr(_f32c(fine)), fine.shape[1], fine.shape[0],
                                        _lib.ptr(out), _s(fine)))
    return out


def point_to_node_partition(points, nodes, point_limit=128, int64=False):
    """pointcloud_partition.py:61-107 -> (point_to_node i32 [N], node_masks bool [M],
    node_knn_indices [M,K], node_knn_masks bool [M,K])."""
    n, m = points.shape[0], nodes.shape[0]
    dev = points.device
    owner = torch.empty(n, dtype=torch.int32, device=dev)
    node_mask = torch.empty(m, dtype=torch.uint8, device=dev)
    knn = torch.empty((m, point_limit), dtype=torch.int64 if int64 else torch.int32, device=dev)
    knn_mask = torch.empty((m, point_limit), dtype=torch.uint8, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    L = _L()
    ws_bytes = L.lcr_point_to_node_ws_bytes(n, m)
    ws = _lib.workspace.get(ws_bytes, dev, slot=4)
    _lib.check(L.lcr_point_to_node(_lib.ptr(_f32c(points)), n, _lib.ptr(_f32c(nodes)), m, point_limit,
                                   _lib.ptr(owner), _lib.ptr(node_mask), _lib.ptr(knn), 1 if int64 else 0,
                                   _lib.ptr(knn_mask), _lib.ptr(status), _lib.ptr(ws), ws.numel(), _s(points)))
    return owner, node_mask.bool(), knn, knn_mask.bool(), status


def scratch(shape, device, slot):
    """float32 tensor of ``shape`` carved from the grow-only per-(device, stream) workspace ``slot``: for the two
    ~1 GB temporaries of the registration tail (patch scores, point-level transport plans of ~17 k patch pairs).
    Asking the caching allocator for them every step made it split and re-grow its largest blocks for more than ten
    forwards (sporadic 40-300 ms cudaMalloc stalls inside a step)."""
    n = 1
    for d in shape:
        n *= int(d)
    buf = _lib.workspace.get(4 * max(n, 1), device, slot=slot)
    return buf[:4 * n].view(torch.float32).view(*shape)


def sinkhorn(scores, row_masks, col_masks, alpha, iters=100, out=None):
    """learnable_sinkhorn.py:13-66: scores [B,M,N] -> [B,M+1,N+1]."""
    b, m, n = scores.shape
    if out is None:
        out = torch.empty((b, m + 1, n + 1), dtype=torch.float32, device=scores.device)
    assert out.shape == (b, m + 1, n + 1) and out.is_contiguous()
    rm = None if row_masks is None else row_masks.to(torch.uint8).contiguous()
    cm = None if col_masks is None else col_masks.to(torch.uint8).contiguous()
    _lib.check(_L().lcr_sinkhorn(_lib.ptr(_f32c(scores)), b, m, n, _lib.ptr(rm), _lib.ptr(cm),
                                 _lib.ptr(alpha.detach().reshape(1)), iters, _lib.ptr(out), _s(scores)))
    return out


def coarse_matching(log_scores, defer=False):
    """superpoint_matching.py:129-160 -> (ref_idx i32 [P], src_idx i32 [P], scores f32 [P]).
    defer=True returns capacity-sized arrays and the device count (no host sync)."""
    r, c = log_scores.shape[0] - 1, log_scores.shape[1] - 1
    dev = log_scores.device
    cap = r + c
    oi = torch.empty(cap, dtype=torch.int32, device=dev)
    oj = torch.empty(cap, dtype=torch.int32, device=dev)
    os_ = torch.empty(cap, dtype=torch.float32, device=dev)
    cnt = torch.zeros(1, dtype=torch.int32, device=dev)
    L = _L()
    ws_bytes = L.lcr_coarse_matching_ws_bytes(r, c)
    ws = _lib.workspace.get(ws_bytes, dev, slot=5)
    _lib.check(L.lcr_coarse_matching(_lib.ptr(_f32c(log_scores)), r, c, _lib.ptr(oi), _lib.ptr(oj), _lib.ptr(os_),
                                     _lib.ptr(cnt), _lib.ptr(ws), ws.numel(), _s(log_scores)))
    if defer:
        return oi, oj, os_, cnt
    p = int(cnt)            # D2H: the number of node correspondences sizes the dense stage
    return oi[:p], oj[:p], os_[:p]


def patch_scores(feats_a, knn_a, node_a, feats_b, knn_b, node_b, out=None):
    p = node_a.shape[0]
    if out is None:
        out = torch.empty((p, 128, 128), dtype=torch.float32, device=feats_a.device)
    assert out.shape == (p, 128, 128) and out.is_contiguous()
    _lib.check(_L().lcr_patch_scores(_lib.ptr(_f32c(feats_a)), feats_a.shape[0], _lib.ptr(knn_a), _lib.ptr(node_a),
                                     _lib.ptr(_f32c(feats_b)), feats_b.shape[0], _lib.ptr(knn_b), _lib.ptr(node_b), p,
                                     128, feats_a.shape[1], _lib.ptr(out), _s(feats_a)))
    return out


def fine_correspondences(log_scores, knn_mask_a, node_a, knn_mask_b, node_b):
    """-> dict of device arrays (capacity P*256) + pair_off i32 [P+1] (pair_off[P] = total)."""
    p = log_scores.shape[0]
    dev = log_scores.device
    cap = max(p * 256, 1)
    r = {'pair_cnt': torch.empty(max(p, 1), dtype=torch.int32, device=dev),
         'pair_off': torch.zeros(p + 1, dtype=torch.int32, device=dev),
         'pair': torch.empty(cap, dtype=torch.int32, device=dev), 'i': torch.empty(cap, dtype=torch.int32, device=dev),
         'j': torch.empty(cap, dtype=torch.int32, device=dev), 'score': torch.empty(cap, dtype=torch.float32, device=dev)}
    ma, mb = knn_mask_a.to(torch.uint8).contiguous(), knn_mask_b.to(torch.uint8).contiguous()
    ws = torch.empty(max(int(_L().lcr_fine_correspondences_ws_bytes(p)), 1), dtype=torch.uint8, device=dev)
    _lib.check(_L().lcr_fine_correspondences(_lib.ptr(_f32c(log_scores)), p, _lib.ptr(ma), _lib.ptr(node_a),
                                             _lib.ptr(mb), _lib.ptr(node_b), _lib.ptr(r['pair_cnt']),
                                             _lib.ptr(r['pair_off']), _lib.ptr(r['pair']), _lib.ptr(r['i']),
                                             _lib.ptr(r['j']), _lib.ptr(r['score']), _lib.ptr(ws), ws.numel(),
                                             _s(log_scores)))
    return r


def corr_points(corr, pts_a, knn_a, node_a, pts_b, knn_b, node_b):
    cap = corr['pair'].shape[0]
    ref = torch.zeros((cap, 3), dtype=torch.float32, device=pts_a.device)
    src = torch.zeros((cap, 3), dtype=torch.float32, device=pts_a.device)
    p = node_a.shape[0]
    _lib.check(_L().lcr_corr_points(_lib.ptr(corr['pair']), _lib.ptr(corr['i']), _lib.ptr(corr['j']),
                                    _lib.ptr(corr['pair_off'][p:]), cap, _lib.ptr(_f32c(pts_a)), _lib.ptr(knn_a),
                                    _lib.ptr(node_a), _lib.ptr(_f32c(pts_b)), _lib.ptr(knn_b), _lib.ptr(node_b),
                                    _lib.ptr(ref), _lib.ptr(src), _s(pts_a)))
    return ref, src


def local_global_registration(ref, src, scores, pair_off, radius=0.45, min_corr=3, steps=5):
    """local_global_registration.py:140-202 on the device -> 4x4 transform (src -> ref)."""
    p = pair_off.shape[0] - 1
    cap = ref.shape[0]
    T = torch.zeros((4, 4), dtype=torch.float32, device=ref.device)
    L = _L()
    ws_bytes = L.lcr_lgr_ws_bytes(p, cap)
    ws = _lib.workspace.get(ws_bytes, ref.device, slot=6)
    _lib.check(L.lcr_local_global_registration(_lib.ptr(_f32c(ref)), _lib.ptr(_f32c(src)), _lib.ptr(_f32c(scores)),
                                               _lib.ptr(pair_off), p, cap, float(radius), min_corr, steps,
                                               _lib.ptr(T), _lib.ptr(ws), ws.numel(), _s(ref)))
    return T


# ------------------------------------------------------------------ batched matching tail (all pairs of a chunk)
def point_to_node_batched(points, pts_stacks, nodes, node_stacks, point_limit=128, want_local=True):
    """pointcloud_partition.py:61-107 for every cloud of the batch at once.  ``pts_stacks`` / ``node_stacks``:
    ops.Stacks over the clouds.  -> (node_mask u8 [M], knn_global i32 [M,K] (global point rows, pad = N),
    knn_local i64 [M,K] (rows local to the cloud, pad = its size: the reference's table), knn_mask u8 [M,K],
    status i32 [1])."""
    n, m = points.shape[0], nodes.shape[0]
    dev = points.device
    node_mask = torch.empty(m, dtype=torch.uint8, device=dev)
    knn_g = torch.empty((m, point_limit), dtype=torch.int32, device=dev)
    knn_l = torch.empty((m, point_limit), dtype=torch.int64, device=dev) if want_local else None
    knn_mask = torch.empty((m, point_limit), dtype=torch.uint8, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    L = _L()
    ws = _lib.workspace.get(L.lcr_point_to_node_ws_bytes(n, m), dev, slot=4)
    _lib.check(L.lcr_point_to_node_batched(_lib.ptr(_f32c(points)), n, _lib.ptr(pts_stacks.off), _lib.ptr(_f32c(nodes)), m,
                                           _lib.ptr(node_stacks.off), pts_stacks.n, pts_stacks.max_rows,
                                           max(node_stacks.max_rows, 1), point_limit, None, _lib.ptr(node_mask),
                                           _lib.ptr(knn_g), _lib.ptr(knn_l), _lib.ptr(knn_mask), _lib.ptr(status),
                                           _lib.ptr(ws), ws.numel(), _s(points)))
    return node_mask, knn_g, knn_l, knn_mask, status


def node_scores(feats, node_stacks, node_mask, n_pairs, m_max, n_max):
    """LCRNet.py:196-199 for all pairs: [P, m_max, n_max] zero-padded scores + padded u8 masks."""
    dev = feats.device
    out = torch.empty((n_pairs, m_max, n_max), dtype=torch.float32, device=dev)
    rm = torch.empty((n_pairs, m_max), dtype=torch.uint8, device=dev)
    cm = torch.empty((n_pairs, n_max), dtype=torch.uint8, device=dev)
    _lib.check(_L().lcr_node_scores(_lib.ptr(_f32c(feats)), feats.shape[1], _lib.ptr(node_stacks.off), _lib.ptr(node_mask),
                                    n_pairs, m_max, n_max, _lib.ptr(out), _lib.ptr(rm), _lib.ptr(cm), _s(feats)))
    return out, rm, cm


def coarse_matching_batched(log_scores):
    """superpoint_matching.py:129-160 on [P, R+1, C+1] -> capacity arrays [P, R+C] (i, j, score) and counts i32 [P]."""
    p, r, c = log_scores.shape[0], log_scores.shape[1] - 1, log_scores.shape[2] - 1
    dev = log_scores.device
    cap = r + c
    oi = torch.empty((p, cap), dtype=torch.int32, device=dev)
    oj = torch.empty((p, cap), dtype=torch.int32, device=dev)
    os_ = torch.empty((p, cap), dtype=torch.float32, device=dev)
    cnt = torch.zeros(p, dtype=torch.int32, device=dev)
    ws_bytes = 4 * p * (2 * (r + 1) + (c + 1)) + 4096
    ws = _lib.workspace.get(ws_bytes, dev, slot=5)
    _lib.check(_L().lcr_coarse_matching_batched(_lib.ptr(_f32c(log_scores)), p, r, c, _lib.ptr(oi), _lib.ptr(oj),
                                                _lib.ptr(os_), _lib.ptr(cnt), _lib.ptr(ws), ws.numel(), _s(log_scores)))
    return oi, oj, os_, cnt


def gather_coarse(oi, oj, os_, patch_off, node_stacks, total):
    """One patch list for the batch: (ci_global, cj_global, ci_local, cj_local, scores, patch_pair), each [total]."""
    dev = oi.device
    p = oi.shape[0]
    i32 = lambda: torch.empty(max(total, 1), dtype=torch.int32, device=dev)
    ci_g, cj_g, ci_l, cj_l, pp = i32(), i32(), i32(), i32(), i32()
    cs = torch.empty(max(total, 1), dtype=torch.float32, device=dev)
    _lib.check(_L().lcr_gather_coarse(_lib.ptr(oi), _lib.ptr(oj), _lib.ptr(os_), oi.shape[1], _lib.ptr(patch_off),
                                      _lib.ptr(node_stacks.off), p, _lib.ptr(ci_g), _lib.ptr(cj_g), _lib.ptr(ci_l),
                                      _lib.ptr(cj_l), _lib.ptr(cs), _lib.ptr(pp), _s(oi)))
    return ci_g[:total], cj_g[:total], ci_l[:total], cj_l[:total], cs[:total], pp[:total]


def lgr_batched(ref, src, scores, c_pair, pair_off, patch_pair, patch_off, n_scan_pairs, radius=0.45, min_corr=3,
                steps=5):
    """local_global_registration.py:140-202 for all scan pairs at once -> [S, 4, 4]."""
    t = pair_off.shape[0] - 1
    cap = ref.shape[0]
    T = torch.zeros((n_scan_pairs, 4, 4), dtype=torch.float32, device=ref.device)
    L = _L()
    ws = _lib.workspace.get(L.lcr_lgr_batched_ws_bytes(t, n_scan_pairs, cap), ref.device, slot=6)
    _lib.check(L.lcr_lgr_batched(_lib.ptr(_f32c(ref)), _lib.ptr(_f32c(src)), _lib.ptr(_f32c(scores)), _lib.ptr(c_pair),
                                 _lib.ptr(pair_off), t, _lib.ptr(patch_pair), _lib.ptr(patch_off), n_scan_pairs, cap,
                                 float(radius), min_corr, steps, _lib.ptr(T), _lib.ptr(ws), ws.numel(), _s(ref)))
    return T
