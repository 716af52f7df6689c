import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from lcrnet_b200 import ops
def t(db, k=25):
    for _ in range(2): ops.l2_topk(db, db, k)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); ops.l2_topk(db, db, k); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)
g = torch.Generator().manual_seed(0)
rnd = torch.nn.functional.normalize(torch.randn(4000, 256, generator=g), dim=1).cuda()
print('random 4000x4000: %.3f ms' % t(rnd))
base = torch.nn.functional.normalize(torch.randn(64, 256, generator=g), dim=1)
for noise in (1e-2, 1e-3, 1e-4, 0.0):
    db = torch.nn.functional.normalize(base.repeat(63, 1)[:4000] + noise * torch.randn(4000, 256, generator=g), dim=1).cuda()
    print('64 distinct x 63 copies, noise %g: %.3f ms' % (noise, t(db)))
# copies ordered so that later rows are closer (worst case for insertion: every candidate beats the k-th)
q = torch.nn.functional.normalize(torch.randn(1, 256, generator=g), dim=1)
db = torch.nn.functional.normalize(q + torch.linspace(1.0, 0.0, 4000)[:, None] * torch.randn(4000, 256, generator=g), dim=1).cuda()
print('monotonically improving candidates: %.3f ms' % t(db))
