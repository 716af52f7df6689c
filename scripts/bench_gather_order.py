"""Experiment: KPConv gather time with the queries processed in hash order (the reference's point order)
vs. spatially grouped order (sorted by grid cell).  Usage: python scripts/bench_gather_order.py n_scans level c_in"""
import ctypes
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
import lcrnet_b200  # noqa
from lcrnet_b200 import _lib, checkpoint, ops, synth
from lcrnet_b200 import data as gdata

n_scans, level, c_in = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
scans = [synth.make_scan(i // 2, 7351 + i) for i in range(n_scans)]
lim = [57, 58, 59, 54]
d = gdata.scans_collate_fn_stack_mode(scans, 4, 0.3, 1.275, lim, pre_voxel=0.3)
p, nb = d['points'][level], d['neighbors'][level]
lens = d['lengths'][level]
rng = np.random.default_rng(0)
feats = torch.from_numpy(rng.standard_normal((p.shape[0], c_in)).astype(np.float32)).cuda()
w = torch.from_numpy((rng.standard_normal((15, c_in, c_in)) * 0.1).astype(np.float32)).cuda()
kp = torch.from_numpy(checkpoint.default_kernel_points(1.275 * 2 ** level, rng))
w_nk = w.reshape(-1, c_in).t().contiguous()
L = _lib.lib()
cell = 1.275 * 2 ** level
scan_id = torch.repeat_interleave(torch.arange(len(lens), device=p.device), lens.to(p.device))
c = torch.floor((p - p.min(0).values) / cell).long()
key = ((scan_id * 4096 + c[:, 2]) * 4096 + c[:, 1]) * 4096 + c[:, 0]
perm = torch.argsort(key)
for name, order in (('hash-order', None), ('cell-order', perm)):
    q = p if order is None else p[order].contiguous()
    t = nb if order is None else nb[order].contiguous()
    for mode, mname in ((1, 'dense'), (5, 'packed')):
        L.lcr_set_gather_mode(mode)
        for it in range(3):
            L.lcr_profile_begin()
            ops.kpconv(feats, q, p, t, kp.cuda(), 0.6 * 2 ** level, w, None, weights_nk=w_nk, kernel_points_host=kp)
            torch.cuda.synchronize()
            n = L.lcr_profile_end()
        for i in range(n):
            nm = ctypes.create_string_buffer(64)
            ms, fl, by = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
            L.lcr_profile_get(i, nm, 64, ctypes.byref(ms), ctypes.byref(fl), ctypes.byref(by))
            if nm.value.decode() == 'kpconv_gather':
                print('%-10s %-6s rows %d C %d  %.3f ms' % (name, mname, p.shape[0], c_in, ms.value))
