"""Experiment: does cutting a KPConv (gather -> wf -> contraction) into row chunks whose wf block stays in L2 pay?
Calls the C entry point per chunk with offset query pointers and ONE reused workspace."""
import os
import sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from lcrnet_b200 import _lib, ops, synth
from lcrnet_b200 import data as gdata

scans = []
for i in range(16):
    ref, src, _ = synth.make_pair(i, 7351 + i)
    scans += [ref, src]
limits = [57, 58, 59, 54]
d = gdata.scans_collate_fn_stack_mode(scans, 4, 0.3, 1.275, limits, pre_voxel=0.3, stack_size=1, int32=True)
L = _lib.lib()
g = torch.Generator().manual_seed(0)
kp = (torch.rand(15, 3, generator=g) - 0.5)
for level, c_in, c_out in ((0, 32, 32), (0, 32, 64), (1, 64, 64), (1, 64, 128), (1, 128, 128)):
    pts = d['points'][level].contiguous()
    idx = ops.as_index32(d['neighbors'][level]).contiguous()
    n = pts.shape[0]
    sigma = 0.3 * 2 ** level * 2.0
    kpl = (kp * sigma).cuda().contiguous()
    kph = kpl.cpu().contiguous()
    feats = torch.randn(n, c_in, generator=g).cuda()
    flags = ops.row_flags(feats)
    w = (torch.randn(15, c_in, c_out, generator=g) / (15 * c_in) ** 0.5).cuda()
    w_nk = w.reshape(15 * c_in, c_out).t().contiguous()
    w_hi, w_lo = ops.tf32_split(w_nk)
    bias = torch.zeros(c_out).cuda()
    out = torch.empty(n, c_out).cuda()
    ws_bytes = L.lcr_kpconv_ws_bytes2(n, n, c_in)
    ws = torch.empty(ws_bytes, dtype=torch.uint8).cuda()
    sp = _lib.stream_ptr(pts.device)

    def run(chunk):
        r0 = 0
        while r0 < n:
            m = min(chunk, n - r0)
            _lib.check(L.lcr_kpconv(_lib.ptr(feats), _lib.ptr(flags), n, pts.data_ptr() + 12 * r0, m, _lib.ptr(pts),
                                    idx.data_ptr() + 4 * idx.stride(0) * r0, idx.stride(0), idx.shape[1], _lib.ptr(kpl),
                                    kph.data_ptr(), float(sigma), _lib.ptr(w), _lib.ptr(w_hi), _lib.ptr(w_lo),
                                    _lib.ptr(bias), c_in, c_out, out.data_ptr() + 4 * c_out * r0, _lib.ptr(ws), ws.numel(),
                                    sp))
            r0 += m

    res = []
    ref = None
    for chunk in (n, 148 * 128 * 4, 148 * 128 * 2, 148 * 128, 74 * 128):
        for _ in range(2):
            run(chunk)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            run(chunk)
        b.record()
        torch.cuda.synchronize()
        if ref is None:
            ref = out.clone()
        res.append('%d rows (%.0f MB wf): %.3f ms%s' % (chunk, chunk * 60.0 * c_in / 1e6, a.elapsed_time(b) / 3,
                                                     '' if torch.equal(ref, out) else ' MISMATCH'))
    print('level %d n=%d c_in=%d c_out=%d: ' % (level, n, c_in, c_out) + '; '.join(res), flush=True)
