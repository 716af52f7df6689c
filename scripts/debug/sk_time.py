"""Diagnostic: point-level Sinkhorn time vs iteration count (fixed cost / per-iteration cost)."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from lcrnet_b200 import pair_ops as P
n = 148 * 8
g = torch.Generator().manual_seed(0)
s = (torch.randn(n, 128, 128, generator=g) * 2).cuda()
rm = (torch.rand(n, 128, generator=g) > 0.1).cuda()
cm = (torch.rand(n, 128, generator=g) > 0.1).cuda()
alpha = torch.tensor(0.7).cuda()
for iters in (0, 1, 2, 3, 10, 50, 100, 200):
    for _ in range(2):
        P.sinkhorn(s, rm, cm, alpha, iters)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        P.sinkhorn(s, rm, cm, alpha, iters)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    print('iters %3d: %.3f ms for %d problems = %.1f us per problem-slot (8 waves)' % (iters, ms, n, ms * 1e3 / 8))
