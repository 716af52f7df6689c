"""Diagnostic: is the registration forward bitwise repeatable?  Runs the same batch several times through
PairPipeline and compares every output tensor of every pair with the first run."""
import os
import sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from lcrnet_b200 import checkpoint, lcrnet, pipeline, synth

n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 8
scans = []
for i in range(n_pairs):
    ref, src, _ = synth.make_pair(i, 7351 + i)
    scans += [ref, src]
limits = [57, 58, 59, 54]
net = lcrnet.create_model(lcrnet.default_cfg(limits)).eval()
net.load_state_dict(checkpoint.random_state_dict('lcrnet', 7351), strict=True)
net = net.cuda()
net.keep_intermediates = len(sys.argv) > 2 and sys.argv[2] == "keep"
pipe = pipeline.PairPipeline(net, limits, 4, 0.3, 1.275, pre_voxel=0.3, n_streams=1)
dev = torch.from_numpy(np.concatenate(scans, 0)).cuda()
lens = [len(s) for s in scans]


def snapshot():
    outs = pipe(dev, lens)
    torch.cuda.synchronize()
    return [{k: (v.clone() if torch.is_tensor(v) else v) for k, v in o.items()} for o in outs]


first = snapshot()
bad = {}
for r in range(6):
    cur = snapshot()
    for p, (a, b) in enumerate(zip(first, cur)):
        for k in a:
            x, y = a[k], b[k]
            if isinstance(x, tuple):
                x, y = x[0], y[0]
            if torch.is_tensor(x) and (x.shape != y.shape or not torch.equal(x, y)):
                bad.setdefault(k, set()).add(p)
print('keys that differ between runs:', {k: sorted(v)[:8] for k, v in bad.items()} or 'none')
