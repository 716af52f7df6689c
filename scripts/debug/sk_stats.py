"""Diagnostic: Sinkhorn iteration statistics (LOG / LIN / discarded / absorptions per problem) on real pairs."""
import sys, os, ctypes
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from lcrnet_b200 import synth, checkpoint, lcrnet, _lib, pair_ops as P
from lcrnet_b200 import data as gdata
scans = []
for i in range(4):
    ref, src, _ = synth.make_pair(i, 7351 + i)
    scans += [ref, src]
limits = gdata.calibrate_neighbors_scans(scans[:2], 4, 0.3, 1.275, pre_voxel=0.3, scans_per_sample=2)
net = lcrnet.create_model(lcrnet.default_cfg(limits)).eval()
net.load_state_dict(checkpoint.random_state_dict('lcrnet', 7351), strict=True)
net = net.cuda()
orig = P.sinkhorn
def wrapped(scores, rm, cm, alpha, iters=100, out=None):
    L = _lib.lib()
    st = (ctypes.c_int64 * 4)()
    L.lcr_sinkhorn_stats(st, 1)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out = orig(scores, rm, cm, alpha, iters, out=out)
    b.record()
    torch.cuda.synchronize()
    L.lcr_sinkhorn_stats(st, 1)
    n = scores.shape[0]
    print('sinkhorn %s: %.3f ms; per problem LOG %.1f LIN %.1f discarded %.1f absorptions %.1f; score range %.1f .. %.1f' % (
        tuple(scores.shape), a.elapsed_time(b), st[0] / n, st[1] / n, st[2] / n, st[3] / n, float(scores.min()), float(scores.max())))
    return out
P.sinkhorn = wrapped
d = gdata.scans_collate_fn_stack_mode(scans, 4, 0.3, 1.275, limits, pre_voxel=0.3, stack_size=2, int32=True, upsampling=True)
for _ in range(2):
    out = net(d)
torch.cuda.synchronize()
