"""Diagnostic: where does the HOST time of one descriptor chunk (32 scans) go?  cProfile of collate + encoder +
NetVLAD with the GPU kept busy (no synchronisation inside the profiled region except the operators' own)."""
import cProfile
import os
import pstats
import sys
import time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from lcrnet_b200 import checkpoint, model, ops, synth
from lcrnet_b200 import data as gdata

scans = []
for i in range(16):
    ref, src, _ = synth.make_pair(i, 7351 + i)
    scans += [ref, src]
limits = [57, 58, 59, 54]
net = model.create_model(model.default_cfg()).eval()
net.load_state_dict(checkpoint.random_state_dict('global_descriptor', 7351), strict=True)
net = net.cuda()
ops.prepare(net)
pts = torch.from_numpy(np.concatenate(scans, 0)).cuda()
lens = torch.tensor([len(s) for s in scans], dtype=torch.int64).cuda()


def chunk():
    d = gdata.device_collate(pts, lens, 4, 0.3, 1.275, limits, pre_voxel=0.3, stack_size=1, int32=True)
    return net(d)['anc_global']


for _ in range(3):
    chunk()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    chunk()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print('host time per chunk %.2f ms (queueing only), %.2f ms incl. final sync' % ((t1 - t0) / 5 * 1e3, (t2 - t0) / 5 * 1e3))
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    chunk()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats('tottime').print_stats(28)
