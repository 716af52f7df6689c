"""GPU: the tcgen05 3xTF32 GEMM (gemm_tc.cu) against an fp64 reference -- fp32-class accuracy."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('m,n,k', [(300, 128, 64), (1000, 32, 480), (257, 64, 960), (128, 384, 128),
                                   (5000, 1024, 256), (77, 96, 32), (4096, 256, 3840)])
def test_linear_tc_matches_fp64(m, n, k):
    from lcrnet_b200 import pair_ops as P
    g = torch.Generator().manual_seed(m + n + k)
    x = torch.randn(m, k, generator=g)
    w = torch.randn(n, k, generator=g) * 0.1
    b = torch.randn(n, generator=g)
    rs = torch.rand(m, generator=g) + 0.5
    ref = (x.double() @ w.double().t()) * rs.double()[:, None] + b.double()
    got = P.linear_tc(x.cuda(), w.cuda(), b.cuda(), rowscale=rs.cuda()).cpu().double()
    scale = float(ref.abs().max())
    err = float((got - ref).abs().max()) / scale
    # plain fp32 accumulation over K terms gives ~1e-6; single-pass TF32 would give ~1e-3
    print('tf32x3 max error / max|ref| = %.2e (m=%d n=%d k=%d)' % (err, m, n, k))
    assert err < 2e-5, err
    got_relu = P.linear_tc(x.cuda(), w.cuda(), b.cuda(), relu=True).cpu().double()
    ref_relu = torch.relu(x.double() @ w.double().t() + b.double())
    assert float((got_relu - ref_relu).abs().max()) / scale < 2e-5


def test_linear_tc_strided_views():
    from lcrnet_b200 import pair_ops as P
    g = torch.Generator().manual_seed(0)
    big = torch.randn(200, 384, generator=g).cuda()
    w = torch.randn(128, 128, generator=g).cuda()
    x = big[:, 128:256]                                   # row stride 384
    out = torch.zeros(200, 256).cuda()
    P.linear_tc(x, w, None, out=out[:, 128:])
    ref = x.double() @ w.double().t()
    assert float((out[:, 128:].double() - ref).abs().max()) < 1e-4
    assert float(out[:, :128].abs().max()) == 0.0


@pytest.mark.parametrize('stack_rows,n,k', [([100, 37, 300, 64], 128, 64), ([1000, 33, 999], 32, 480),
                                            ([129, 127, 128], 64, 960), ([450, 350], 512, 128),
                                            ([32, 32, 33, 31 + 32], 256, 1920)])
def test_fused_group_norm_statistics(stack_rows, n, k):
    """GroupNorm statistics from the GEMM epilogue (gemm_tc.cu GnFuse + gn_finalize_blocks_kernel) against an
    fp64 reference, with stack boundaries that cut through the 32-row blocks; the output must equal the
    unfused GEMM bit for bit."""
    from lcrnet_b200 import ops
    g = torch.Generator().manual_seed(sum(stack_rows) + n + k)
    m = sum(stack_rows)
    x = (torch.randn(m, k, generator=g) + 0.3).cuda()
    w = (torch.randn(n, k, generator=g) * 0.1).cuda()
    b = torch.randn(n, generator=g).cuda()
    stacks = ops.Stacks(stack_rows, x.device)
    assert ops.gn_fusable(stacks)
    out, stats = ops.linear(x, w.t().contiguous(), b, w, gn=(stacks, 1e-5, 32))
    plain = ops.linear(x, w.t().contiguous(), b, w)
    assert torch.equal(out, plain)
    ref = x.double() @ w.double().t() + b.double()
    off = np.cumsum([0] + stack_rows)
    for s in range(len(stack_rows)):
        blk = ref[off[s]:off[s + 1]].reshape(stack_rows[s], 32, n // 32)
        mean = blk.mean(dim=(0, 2))
        var = blk.var(dim=(0, 2), unbiased=False)
        rstd = 1.0 / torch.sqrt(var + 1e-5)
        got = stats[s].cpu().double()
        assert float((got[:, 0] - mean.cpu()).abs().max()) < 1e-5 * (1 + float(mean.abs().max()))
        assert float(((got[:, 1] - rstd.cpu()) / rstd.cpu()).abs().max()) < 1e-5
    sep = ops.group_norm_stats(out, stacks, 1e-5, 32)
    assert float((sep - stats).abs().max()) < 1e-5 * float(sep.abs().max())
