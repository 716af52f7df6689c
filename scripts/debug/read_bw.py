import torch
x = torch.randn(1 << 30, device='cuda')        # 4 GiB
for name, fn in (('sum (pure read)', lambda: x.sum()), ('max (pure read)', lambda: x.max()),
                 ('copy (read+write)', lambda: x.clone()), ('fill (pure write)', lambda: x.fill_(1.0))):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    by = x.numel() * 4 * (2 if 'copy' in name else 1)
    print('%-20s %.3f ms  %.0f GB/s' % (name, best, by / best / 1e6))
