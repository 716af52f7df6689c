"""Times the tensor-core GEMM (lcr_linear_tc) on the encoder's shapes at the bench batch (64 scans):
per shape ms, fp32-equivalent TFLOP/s, TF32-MMA TFLOP/s (x3) and the A+C HBM stream in GB/s.
LCR_GEMM_WS selects the kernel (1: operands in shared memory, 3: A operand in tensor memory)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lcrnet_b200 import pair_ops as P
SHAPES = [  # (M, K, N, what)
    (909000, 480, 32, 'KPConv 32->32 @L0'), (361000, 480, 32, 'KPConv 32->32 strided'),
    (361000, 960, 64, 'KPConv 64->64 @L1'), (130000, 960, 64, 'KPConv 64->64 strided'),
    (130000, 1920, 128, 'KPConv 128->128 @L2'), (46000, 1920, 128, 'KPConv 128->128 strided'),
    (46000, 3840, 256, 'KPConv 256->256 @L3'),
    (361000, 256, 64, 'unary 256->64'), (130000, 512, 128, 'unary 512->128'), (46000, 1024, 256, 'unary 1024->256'),
    (46000, 256, 1024, 'unary 256->1024'), (130000, 256, 512, 'unary 256->512'), (130000, 128, 512, 'unary 128->512'),
    (909000, 64, 32, 'unary 64->32'), (909000, 32, 128, 'unary 32->128'),
    (909000, 64, 128, 'unary 64->128'), (361000, 128, 256, 'unary 128->256'), (361000, 64, 256, 'unary 64->256'),
    (361000, 128, 64, 'unary 128->64'), (130000, 128, 512, 'unary 128->512 (2 n-tiles)'),
]
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
tot = 0.0
for m, k, n, what in SHAPES:
    x = torch.randn(m, k, device='cuda')
    w = torch.randn(n, k, device='cuda') * 0.05
    b = torch.zeros(n, device='cuda')
    for _ in range(2):
        P.linear_tc(x, w, b)
    ts = []
    for _ in range(5):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        P.linear_tc(x, w, b)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    tot += ms
    fl = 2.0 * m * k * n
    by = 4.0 * (m * k + m * n)
    print('%-26s M=%7d K=%4d N=%4d  %7.3f ms  %6.1f TFLOP/s fp32-eq (%6.1f TF32 MMA)  %6.0f GB/s' % (
        what, m, k, n, ms, fl / ms / 1e9, 3 * fl / ms / 1e9, by / ms / 1e6))
    del x
print('sum %.3f ms (WS=%s)' % (tot, os.environ.get('LCR_GEMM_WS', '1')))
