"""CPU restatement (torch fp32, eager) of the LCR-Net registration (scan-pair) path.
TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench CPU baseline); never imported by lcr-net_b200/.

Parity status: PINNED against the UNMODIFIED reference Python: tests/test_oracle_pair.py runs
``experiments.lcrnet.model_family.LCRNet`` live when /root/reference is present and compares every
stage, and checks the committed fixture tests/golden/pair_golden.npz generated from it.

Reference followed (experiments/lcrnet/...):
  ThDRoFormer            modules/thdroformer/thdroformer_linear.py:50-96, Rotary3DPosEmb.py:27-39,
                         rpetransformer.py:41-54 (rotary), :57-108, :111-171, :173-220,
                         vanilla_transformer.py:13-144
  Vote layer / NMS       modules/vote/vote.py:149-183, :13-70
  Vote encoder           backbone4.py:121-220
  point->node partition  modules/ops/pointcloud_partition.py:61-107, pairwise_distance.py:4-31
  Sinkhorn               modules/sinkhorn/learnable_sinkhorn.py:13-66
  coarse matching        modules/geotransformer/superpoint_matching.py:129-160
  decoder                backbone4.py:344-373, modules/kpconv/functional.py:6-22
  fine matching + LGR    modules/geotransformer/local_global_registration.py:49-246,
                         modules/registration/procrustes.py:6-73, modules/ops/transformation.py:7-61
  composition            model_family/LCRNet.py:124-321
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import model_oracle as mo
from . import native


# --------------------------------------------------------------------------- transformer
def _lin(sd, p, x):
    return F.linear(x, sd[p + 'weight'], sd.get(p + 'bias'))


def _ln(sd, p, x):
    return F.layer_norm(x, (x.shape[-1],), sd[p + 'weight'], sd[p + 'bias'], 1e-5)


def rotary(x, theta):
    """rpetransformer.py:41-54.  x [H, N, 32], theta [H, N, 16] -> rotated x."""
    x2 = x.reshape(x.shape[0], x.shape[1], -1, 2)
    rot = torch.stack([-x2[..., 1], x2[..., 0]], -1).reshape(x.shape)
    th = theta.repeat_interleave(2, dim=-1)
    return x * torch.cos(th) + rot * torch.sin(th)


def _heads(x, h):
    return x.reshape(x.shape[0], h, -1).permute(1, 0, 2)          # [N, H*c] -> [H, N, c]


def attention_layer(sd, p, x, mem, theta_x=None, theta_mem=None, heads=4):
    """(RPE)AttentionLayer + AttentionOutput: one transformer layer.  x [N,128], mem [M,128].
    theta given -> rotary self-attention (q and k both rotated with THEIR OWN point's angles)."""
    a = p + 'attention.attention.'
    q, k, v = (_heads(_lin(sd, a + n, t), heads) for n, t in (('proj_q.', x), ('proj_k.', mem), ('proj_v.', mem)))
    if theta_x is not None:
        q = rotary(q, _heads(theta_x, heads))
        k = rotary(k, _heads(theta_mem, heads))
    s = torch.softmax(torch.einsum('hnc,hmc->hnm', q, k) / q.shape[-1] ** 0.5, dim=-1)
    h = torch.einsum('hnm,hmc->hnc', s, v).permute(1, 0, 2).reshape(x.shape[0], -1)
    h = _ln(sd, p + 'attention.norm.', _lin(sd, p + 'attention.linear.', h) + x)
    o = _lin(sd, p + 'output.squeeze.', F.relu(_lin(sd, p + 'output.expand.', h)))
    return _ln(sd, p + 'output.norm.', h + o)


def thdroformer(sd, ref_points, src_points, ref_feats, src_feats, prefix='transformer.', num_layers=4):
    """thdroformer_linear.py:50-96 with blocks ['self','cross'] * 4, parallel=False."""
    emb = lambda pts: _lin(sd, prefix + 'embedding.encoder2.', _lin(sd, prefix + 'embedding.encoder.', pts))
    th0, th1 = emb(ref_points), emb(src_points)
    f0, f1 = _lin(sd, prefix + 'in_proj.', ref_feats), _lin(sd, prefix + 'in_proj.', src_feats)
    for i in range(2 * num_layers):
        p = prefix + 'transformer.layers.%d.' % i
        if i % 2 == 0:
            f0 = attention_layer(sd, p, f0, f0, th0, th0)
            f1 = attention_layer(sd, p, f1, f1, th1, th1)
        else:
            f0 = attention_layer(sd, p, f0, f1)
            f1 = attention_layer(sd, p, f1, f0)        # sequential: src attends to the UPDATED ref
    return _lin(sd, prefix + 'out_proj.', f0), _lin(sd, prefix + 'out_proj.', f1)


# --------------------------------------------------------------------------- vote / NMS
def vote_layer(sd, points, feats, prefix='vote_encoder.vote.', max_range=4.2):
    """vote.py:149-183 (output_feats=False)."""
    x = feats
    for a, b in ((0, 1), (3, 4)):
        x = F.relu(_ln(sd, prefix + 'mlp_modules.%d.' % b, _lin(sd, prefix + 'mlp_modules.%d.' % a, x)))
    off = _lin(sd, prefix + 'ctr_reg.', x)[:, :3]
    dis = torch.norm(off, p=2, dim=1)
    alpha = torch.where(dis > max_range, max_range / dis, torch.tensor(1.0))
    return points + off * alpha[:, None]


def nms(points, lengths, radius=2.4):
    """vote.py:13-70: greedy, first point always kept, nn.PairwiseDistance (eps 1e-6 added to the
    difference), kept iff ALL distances to kept points are > radius."""
    masks, counts, o = [], [], 0
    for n in lengths:
        n = int(n)
        p = points[o:o + n]
        keep = torch.zeros(n, dtype=torch.bool)
        if n:
            keep[0] = True
        for i in range(1, n):
            d = F.pairwise_distance(p[i][None], p[keep], p=2.0, eps=1e-6)
            if bool((d > radius).all()):
                keep[i] = True
        masks.append(keep)
        counts.append(int(keep.sum()))
        o += n
    return torch.cat(masks), counts


def _radius(q, s, ql, sl, r, limit):
    t = native.radius_neighbors(q.numpy(), s.numpy(), np.asarray(ql, np.int64), np.asarray(sl, np.int64), r, limit)
    return torch.from_numpy(t)


def vote_encoder(sd, feats, data, limits, init_radius=1.275, init_sigma=0.6, prefix='vote_encoder.'):
    """backbone4.py:121-220.  Returns dict with shifted points, node centres, node features."""
    pts_c, len_c = data['points'][-1], [int(x) for x in data['lengths'][-1]]
    shifted = vote_layer(sd, pts_c, feats, prefix + 'vote.')
    keep, counts = nms(shifted, len_c)
    nodes0 = shifted[keep]
    idx = _radius(nodes0, shifted, counts, len_c, 2.4, limits[-1])
    pad = shifted.shape[0]
    sp = torch.cat([shifted, torch.zeros(1, 3)], 0)
    valid = idx < pad
    centres = (sp[idx.reshape(-1)].reshape(idx.shape + (3,)) * valid[..., None]).sum(1) / valid.sum(1, keepdim=True)
    sub = _radius(centres, pts_c, counts, len_c, init_radius * 8, limits[-2])
    nb = _radius(centres, centres, counts, counts, init_radius * 16, limits[-1])
    x = mo.residual_block(sd, prefix + 'encoder6_1.', feats, centres, pts_c, sub, init_sigma * 8, True)
    x = mo.residual_block(sd, prefix + 'encoder6_2.', x, centres, centres, nb, init_sigma * 16, False)
    x = mo.residual_block(sd, prefix + 'encoder6_3.', x, centres, centres, nb, init_sigma * 16, False)
    return {'shifted': shifted, 'keep': keep, 'counts': counts, 'centres': centres, 'feats': x,
            'node_knn': idx, 'subsampling': sub, 'neighbors': nb}


# --------------------------------------------------------------------------- grouping
def pairwise_distance(x, y):
    """pairwise_distance.py:4-31 (matmul form, clamp 1e-12)."""
    xy = x @ y.t()
    return ((x ** 2).sum(-1)[:, None] - 2 * xy + (y ** 2).sum(-1)[None, :]).clamp(min=1e-12)


def point_to_node_partition(points, nodes, point_limit=128):
    """pointcloud_partition.py:61-107."""
    d = pairwise_distance(nodes, points)
    owner = d.min(dim=0)[1]
    node_masks = torch.zeros(nodes.shape[0], dtype=torch.bool)
    node_masks[owner] = True
    match = torch.zeros_like(d, dtype=torch.bool)
    match[owner, torch.arange(points.shape[0])] = True
    d = d.masked_fill(~match, 1e12)
    knn = d.topk(k=point_limit, dim=1, largest=False)[1]
    knn_masks = owner[knn] == torch.arange(nodes.shape[0])[:, None]
    knn = knn.masked_fill(~knn_masks, points.shape[0])
    return owner, node_masks, knn, knn_masks


# --------------------------------------------------------------------------- optimal transport
def sinkhorn(scores, row_masks, col_masks, alpha, iters=100, inf=1e12):
    """learnable_sinkhorn.py:13-66.  scores [B,M,N] -> [B,M+1,N+1]."""
    b, m, n = scores.shape
    prm = torch.zeros(b, m + 1, dtype=torch.bool)
    prm[:, :m] = ~row_masks
    pcm = torch.zeros(b, n + 1, dtype=torch.bool)
    pcm[:, :n] = ~col_masks
    s = torch.cat([torch.cat([scores, alpha.expand(b, m, 1)], -1), alpha.expand(b, 1, n + 1)], 1)
    s = s.masked_fill(prm[:, :, None] | pcm[:, None, :], -inf)
    nvr, nvc = row_masks.float().sum(1), col_masks.float().sum(1)
    norm = -torch.log(nvr + nvc)
    log_mu = norm[:, None].repeat(1, m + 1)
    log_mu[:, m] = torch.log(nvc) + norm
    log_mu[prm] = -inf
    log_nu = norm[:, None].repeat(1, n + 1)
    log_nu[:, n] = torch.log(nvr) + norm
    log_nu[pcm] = -inf
    u, v = torch.zeros_like(log_mu), torch.zeros_like(log_nu)
    for _ in range(iters):
        u = log_mu - torch.logsumexp(s + v[:, None, :], dim=2)
        v = log_nu - torch.logsumexp(s + u[:, :, None], dim=1)
    return s + u[:, :, None] + v[:, None, :] - norm[:, None, None]


def coarse_matching(log_scores):
    """superpoint_matching.py:129-160 (num_correspondences None): (M+1, N+1) log scores."""
    s = torch.exp(log_scores)
    col_best = s.argmax(dim=0)
    src_mat = torch.zeros_like(s)
    src_mat[col_best, torch.arange(s.shape[1])] = s.max(dim=0)[0]
    src_corr = src_mat > s[-1, :][None, :]
    row_best = s.argmax(dim=1)
    ref_mat = torch.zeros_like(s)
    ref_mat[torch.arange(s.shape[0]), row_best] = s.max(dim=1)[0]
    ref_corr = ref_mat > s[:, -1][:, None]
    corr = (ref_corr | src_corr)[:-1, :-1]
    ij = corr.nonzero()
    return ij[:, 0], ij[:, 1], s[ij[:, 0], ij[:, 1]]


# --------------------------------------------------------------------------- decoder
def nearest_upsample(x, up_idx):
    return torch.cat([x, torch.zeros_like(x[:1])], 0)[up_idx[:, 0]]


def kpdecoder(sd, feats_list, data, prefix='kpdecoder.'):
    """backbone4.py:344-373 -> finest-level features [N0, 128]."""
    up = data['upsampling']
    l3 = mo.unary(sd, prefix + 'decoder3.', torch.cat([nearest_upsample(feats_list[3], up[2]), feats_list[2]], 1))
    l2 = mo.unary(sd, prefix + 'decoder2.', torch.cat([nearest_upsample(l3, up[1]), feats_list[1]], 1))
    l1 = torch.cat([nearest_upsample(l2, up[0]), feats_list[0]], 1)
    return F.linear(l1, sd[prefix + 'decoder1.mlp.weight'], sd[prefix + 'decoder1.mlp.bias'])


# --------------------------------------------------------------------------- fine matching + LGR
def weighted_procrustes(src, ref, w, eps=1e-5):
    """procrustes.py:6-73 -> 4x4 (batched if src is [B,N,3])."""
    squeeze = src.ndim == 2
    if squeeze:
        src, ref, w = src[None], ref[None], w[None]
    w = torch.where(w < 0.0, torch.zeros_like(w), w)
    w = (w / (w.sum(1, keepdim=True) + eps))[..., None]
    sc, rc = (src * w).sum(1, keepdim=True), (ref * w).sum(1, keepdim=True)
    H = (src - sc).permute(0, 2, 1) @ (w * (ref - rc))
    U, _, V = torch.svd(H)
    Ut = U.transpose(1, 2)
    eye = torch.eye(3)[None].repeat(src.shape[0], 1, 1)
    eye[:, -1, -1] = torch.sign(torch.det(V @ Ut))
    R = V @ eye @ Ut
    t = (rc.permute(0, 2, 1) - R @ sc.permute(0, 2, 1)).squeeze(2)
    T = torch.eye(4)[None].repeat(src.shape[0], 1, 1)
    T[:, :3, :3] = R
    T[:, :3, 3] = t
    return T[0] if squeeze else T


def apply_transform(points, T):
    if T.ndim == 2:
        return points @ T[:3, :3].t() + T[:3, 3]
    return points @ T[:, :3, :3].transpose(-1, -2) + T[:, None, :3, 3]


def fine_correspondences(log_scores, ref_masks, src_masks):
    """local_global_registration.py:49-92 + :234-243 with k=1, mutual=False, use_dustbin=True.
    log_scores [P,129,129] -> corr_mat [P,128,128] (bool), score_mat [P,128,128]."""
    s = torch.exp(log_scores)
    P, R, C = s.shape
    ref_mat = torch.zeros_like(s)
    ref_mat.scatter_(2, s.argmax(2, keepdim=True), s.max(2, keepdim=True)[0])
    ref_corr = ref_mat > s[:, :, -1:]
    src_mat = torch.zeros_like(s)
    src_mat.scatter_(1, s.argmax(1, keepdim=True), s.max(1, keepdim=True)[0])
    src_corr = src_mat > s[:, -1:, :]
    corr = (ref_corr | src_corr)[:, :-1, :-1] & (ref_masks[:, :, None] & src_masks[:, None, :])
    return corr, s[:, :-1, :-1] * corr.float()


def lgr_from_lists(ref_c, src_c, sc, b, radius=0.45, min_corr=3, steps=5):
    """local_global_registration.py:140-202 on flat correspondence lists; ``b`` = patch index of every
    correspondence (non-decreasing): one weighted-Procrustes hypothesis per patch with >= min_corr entries, the
    hypothesis with most inliers over ALL correspondences seeds ``steps`` rounds of inlier re-weighting."""
    bounds = [0] + (torch.nonzero(b[1:] != b[:-1])[:, 0] + 1).tolist() + [b.shape[0]]
    chunks = [(x, y) for x, y in zip(bounds[:-1], bounds[1:]) if y - x >= min_corr]
    if chunks:
        mx = max(y - x for x, y in chunks)
        br, bs, bw = torch.zeros(len(chunks), mx, 3), torch.zeros(len(chunks), mx, 3), torch.zeros(len(chunks), mx)
        for n, (x, y) in enumerate(chunks):
            br[n, :y - x], bs[n, :y - x], bw[n, :y - x] = ref_c[x:y], src_c[x:y], sc[x:y]
        Ts = weighted_procrustes(bs, br, bw)
        res = torch.linalg.norm(ref_c[None] - apply_transform(src_c[None], Ts), dim=2)
        inl = res < radius
        best = inl.sum(1).argmax()
        cur = sc * inl[best].float()
    else:
        T = weighted_procrustes(src_c, ref_c, sc)
        cur = sc * (torch.linalg.norm(ref_c - apply_transform(src_c, T), dim=1) < radius).float()
    T = weighted_procrustes(src_c, ref_c, cur)
    for _ in range(steps - 1):
        cur = sc * (torch.linalg.norm(ref_c - apply_transform(src_c, T), dim=1) < radius).float()
        T = weighted_procrustes(src_c, ref_c, cur)
    return T


def local_global_registration(ref_knn_points, src_knn_points, score_mat, corr_mat, radius=0.45, min_corr=3, steps=5):
    """local_global_registration.py:140-202."""
    b, i, j = torch.nonzero(corr_mat, as_tuple=True)
    ref_c, src_c, sc = ref_knn_points[b, i], src_knn_points[b, j], score_mat[b, i, j]
    return ref_c, src_c, sc, lgr_from_lists(ref_c, src_c, sc, b, radius, min_corr, steps)


# --------------------------------------------------------------------------- composition
def lcrnet_forward(sd, data, limits, stages=False):
    """LCRNet.forward (LCRNet.py:274-321) on ONE pair (stack of 2 clouds: pos = ref first)."""
    n_f = [int(x) for x in data['lengths'][0]]
    n_c = [int(x) for x in data['lengths'][-1]]
    pts_f, pts_c = data['points'][0], data['points'][-1]
    feats_list = mo.kpencoder(sd, torch.ones(pts_f.shape[0], 1), data)
    fc = feats_list[-1]
    pos_fc, anc_fc = fc[:n_c[0]], fc[n_c[0]:]
    e0, e1 = thdroformer(sd, pts_c[:n_c[0]], pts_c[n_c[0]:], pos_fc, anc_fc)
    enhanced = torch.cat([e0, e1], 0)
    vd = vote_encoder(sd, enhanced, data, limits)
    out = {'pos_feature_global': mo.netvlad(sd, pos_fc), 'anc_feature_global': mo.netvlad(sd, anc_fc)}
    m0, m1 = vd['counts']
    pos_nodes, anc_nodes = vd['centres'][:m0], vd['centres'][m0:]
    pos_nf, anc_nf = vd['feats'][:m0], vd['feats'][m0:]
    pos_pf, anc_pf = pts_f[:n_f[0]], pts_f[n_f[0]:]
    _, pos_nm, pos_knn, pos_km = point_to_node_partition(pos_pf, pos_nodes)
    _, anc_nm, anc_knn, anc_km = point_to_node_partition(anc_pf, anc_nodes)
    node_scores = (pos_nf @ anc_nf.t() / pos_nf.shape[1] ** 0.5)[None]
    node_ot = sinkhorn(node_scores, pos_nm[None], anc_nm[None], sd['node_optimal_transport.alpha'])[0]
    ci, cj, cs = coarse_matching(node_ot)
    feats_f = kpdecoder(sd, feats_list[:3] + [enhanced], data)
    pos_ff, anc_ff = feats_f[:n_f[0]], feats_f[n_f[0]:]
    pad = lambda x: torch.cat([x, torch.zeros_like(x[:1])], 0)
    pk, ak = pos_knn[ci], anc_knn[cj]
    pkm, akm = pos_km[ci], anc_km[cj]
    pkp, akp = pad(pos_pf)[pk], pad(anc_pf)[ak]
    pkf, akf = pad(pos_ff)[pk], pad(anc_ff)[ak]
    ms_raw = torch.einsum('bnd,bmd->bnm', pkf, akf) / feats_f.shape[1] ** 0.5
    ms = sinkhorn(ms_raw, pkm, akm, sd['optimal_transport.alpha'])
    corr_mat, score_mat = fine_correspondences(ms, pkm, akm)
    ref_c, src_c, sc, T = local_global_registration(pkp, akp, score_mat, corr_mat)
    out.update({'estimated_transform': T, 'pos_corr_points': ref_c, 'anc_corr_points': src_c, 'corr_scores': sc,
                'pos_node_corr_indices': ci, 'anc_node_corr_indices': cj, 'pos_feats_f': pos_ff, 'anc_feats_f': anc_ff,
                'pos_points_c': pos_nodes, 'anc_points_c': anc_nodes, 'pos_feats_c': pos_nf, 'anc_feats_c': anc_nf,
                'shifted_pos_points_c': vd['shifted'][:n_c[0]], 'shifted_anc_points_c': vd['shifted'][n_c[0]:],
                'length': vd['counts']})
    if stages:
        out['_stages'] = {'feats_c': fc, 'enhanced': enhanced, 'vote': vd, 'node_ot': node_ot, 'feats_f': feats_f,
                          'pos_knn': pos_knn, 'anc_knn': anc_knn, 'pos_node_masks': pos_nm, 'anc_node_masks': anc_nm,
                          'point_ot': ms, 'corr_mat': corr_mat, 'node_scores': node_scores[0], 'point_scores': ms_raw}
    return out
