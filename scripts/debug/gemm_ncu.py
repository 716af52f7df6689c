import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from lcrnet_b200 import pair_ops as P
m, k, n = (int(x) for x in sys.argv[1:4])
x = torch.randn(m, k, device='cuda')
w = torch.randn(n, k, device='cuda') * 0.05
b = torch.zeros(n, device='cuda')
for _ in range(3):
    P.linear_tc(x, w, b)
torch.cuda.synchronize()
