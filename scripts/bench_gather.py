"""Micro-benchmark of the KPConv gather variants on one pyramid level of a synthetic batch.
Usage: python scripts/bench_gather.py [n_scans] [level] [c_in]   (prints ms per variant)"""
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
import lcrnet_b200  # noqa
from lcrnet_b200 import _lib, checkpoint, ops, synth
from lcrnet_b200 import data as gdata

n_scans = int(sys.argv[1]) if len(sys.argv) > 1 else 8
level = int(sys.argv[2]) if len(sys.argv) > 2 else 1
c_in = int(sys.argv[3]) if len(sys.argv) > 3 else 64
scans = [synth.make_scan(i // 2, 7351 + i) for i in range(n_scans)]
lim = [57, 58, 59, 54]
d = gdata.scans_collate_fn_stack_mode(scans, 4, 0.3, 1.275, lim, pre_voxel=0.3)
p, nb = d['points'][level], d['neighbors'][level]
rng = np.random.default_rng(0)
feats = torch.from_numpy(rng.standard_normal((p.shape[0], c_in)).astype(np.float32)).cuda()
w = torch.from_numpy((rng.standard_normal((15, c_in, c_in)) * 0.1).astype(np.float32)).cuda()
kp = torch.from_numpy(checkpoint.default_kernel_points(1.275 * 2 ** level, rng))
w_nk = w.reshape(-1, c_in).t().contiguous()
L = _lib.lib()
outs = []
for mode, name in ((0, 'exact'), (1, 'dense'), (2, 'sparse'), (6, 'mma'), (7, 'group')):
    L.lcr_set_gather_mode(mode)
    for it in range(3):
        L.lcr_profile_begin()
        out = ops.kpconv(feats, p, p, nb, kp.cuda(), 0.6 * 2 ** level, w, None, weights_nk=w_nk, kernel_points_host=kp)
        torch.cuda.synchronize()
        n = L.lcr_profile_end()
    import ctypes
    for i in range(n):
        nm = ctypes.create_string_buffer(64)
        ms, fl, by = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
        L.lcr_profile_get(i, nm, 64, ctypes.byref(ms), ctypes.byref(fl), ctypes.byref(by))
        print('%-7s rows %d H %d C %d  %-14s %.3f ms' % (name, p.shape[0], nb.shape[1], c_in, nm.value.decode(), ms.value))
    outs.append(out)
print('max |dense-exact| %.2e  |sparse-exact| %.2e  |mma-exact| %.2e |group-exact| %.2e' % (float((outs[1] - outs[0]).abs().max()), float((outs[2] - outs[0]).abs().max()), float((outs[3] - outs[0]).abs().max()), float((outs[4] - outs[0]).abs().max())))
