import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lcrnet_b200 import pair_ops as P
def t(fn, n=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
g = torch.Generator().manual_seed(0)
alpha = torch.tensor(1.0).cuda()
for (b, m, n) in [(1, 324, 312), (16, 324, 312), (600, 128, 128), (148 * 3, 128, 128)]:
    s = torch.randn(b, m, n, generator=g).cuda()
    rm = torch.ones(b, m, dtype=torch.bool).cuda(); cm = torch.ones(b, n, dtype=torch.bool).cuda()
    print((b, m, n), '%.3f ms' % t(lambda: P.sinkhorn(s, rm, cm, alpha)))
