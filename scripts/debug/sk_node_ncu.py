"""Diagnostic: node-level Sinkhorn (32 problems of 392 x 384, masked tails) for ncu captures / timing."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from lcrnet_b200 import pair_ops as P
n, m, k = 32, 392, 384
g = torch.Generator().manual_seed(0)
s = (torch.randn(n, m, k, generator=g) * 30).cuda()
rm = (torch.arange(m)[None, :] < torch.randint(330, m + 1, (n, 1), generator=g)).cuda()
cm = (torch.arange(k)[None, :] < torch.randint(330, k + 1, (n, 1), generator=g)).cuda()
alpha = torch.tensor(0.7).cuda()
for _ in range(3):
    P.sinkhorn(s, rm, cm, alpha, 100)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    P.sinkhorn(s, rm, cm, alpha, 100)
b.record()
torch.cuda.synchronize()
print('node sinkhorn %d x %d x %d: %.3f ms per call' % (n, m, k, a.elapsed_time(b) / 5))
