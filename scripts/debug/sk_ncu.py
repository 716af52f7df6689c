import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from lcrnet_b200 import pair_ops as P
n = 148 * 2
g = torch.Generator().manual_seed(0)
s = (torch.randn(n, 128, 128, generator=g) * 2).cuda()
rm = (torch.rand(n, 128, generator=g) > 0.1).cuda()
cm = (torch.rand(n, 128, generator=g) > 0.1).cuda()
alpha = torch.tensor(0.7).cuda()
for _ in range(3):
    P.sinkhorn(s, rm, cm, alpha, 100)
torch.cuda.synchronize()
