"""Diagnostic: node-level Sinkhorn on full-size pairs: score magnitude, GPU kernel on the ORACLE's scores."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from lcrnet_b200 import synth, checkpoint, pair_ops as P
from oracle import native as on, model_oracle as mo, pair_oracle as po
sd = checkpoint.random_state_dict('lcrnet', 7351)
limits = [55, 56, 55, 53]
for scene, seed in [(11, 8001), (12, 8002)]:
    ref, src, _ = synth.make_pair(scene, seed)
    p0, l0 = on.grid_subsample(np.concatenate([ref, src]), np.array([len(ref), len(src)], dtype=np.int64), 0.3)
    data = mo.precompute_pyramid(p0, l0, limits=limits)
    with torch.no_grad():
        out = po.lcrnet_forward(sd, data, limits, stages=True)
    st = out['_stages']
    pos_nf, anc_nf = out['pos_feats_c'], out['anc_feats_c']
    sc = (pos_nf @ anc_nf.t() / pos_nf.shape[1] ** 0.5)[None]
    print('scene', scene, 'nodes', out['length'], 'score range', float(sc.min()), float(sc.max()), 'feat max', float(pos_nf.abs().max()))
    rm, cm = st['pos_node_masks'][None], st['anc_node_masks'][None]
    alpha = sd['node_optimal_transport.alpha']
    ref_ot = st['node_ot']
    ref64 = po.sinkhorn(sc.double(), rm, cm, alpha.double())[0]
    got = P.sinkhorn(sc.cuda().contiguous(), rm.cuda(), cm.cuda(), alpha.cuda())[0].cpu()
    valid = ref_ot > -1e11
    print('  L range', float(ref_ot[valid].min()), float(ref_ot[valid].max()))
    print('  gpu vs oracle %.3e  gpu vs fp64 %.3e  oracle vs fp64 %.3e' % (
        float(((got - ref_ot).abs() * valid).max()), float(((got.double() - ref64).abs() * valid).max()),
        float(((ref_ot.double() - ref64).abs() * valid).max())))
    # padded to a larger problem like the batched path does
    m, n = sc.shape[1], sc.shape[2]
    M, N = m + 13, n + 7
    scp = torch.zeros(1, M, N); scp[0, :m, :n] = sc[0]
    rmp = torch.zeros(1, M, dtype=torch.bool); rmp[0, :m] = rm[0]
    cmp_ = torch.zeros(1, N, dtype=torch.bool); cmp_[0, :n] = cm[0]
    gp = P.sinkhorn(scp.cuda(), rmp.cuda(), cmp_.cuda(), alpha.cuda())[0].cpu()
    rows = list(range(m)) + [M]; cols = list(range(n)) + [N]
    gp = gp[rows][:, cols]
    print('  padded: gpu vs fp64 %.3e' % float(((gp.double() - ref64).abs() * valid).max()))
    for it in (4, 8, 20, 100):
        r = po.sinkhorn(sc, rm, cm, alpha, iters=it)[0]
        g = P.sinkhorn(sc.cuda().contiguous(), rm.cuda(), cm.cuda(), alpha.cuda(), iters=it)[0].cpu()
        print('  iters %3d: gpu vs oracle %.3e' % (it, float(((g - r).abs() * valid).max())))
