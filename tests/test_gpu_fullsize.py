"""GPU: the BASELINE.json sizes (65 536-point scans, 4000-row descriptor databases), checked through the oracle
where it finishes in seconds (one full scan) and through size-independent properties otherwise (unit norm,
scan independence, sortedness, radius bound, self match)."""
import numpy as np
import pytest
import torch

from lcrnet_b200 import checkpoint, synth
from oracle import model_oracle as mo
from oracle import native as on

pytestmark = pytest.mark.gpu

LIMITS = [57, 58, 59, 54]          # the calibrated limits of the bench workload


@pytest.fixture(scope='module')
def sd():
    return checkpoint.random_state_dict('global_descriptor', 7351)


@pytest.fixture(scope='module')
def net(sd):
    from lcrnet_b200 import model
    n = model.create_model(model.default_cfg()).eval()
    n.load_state_dict(sd, strict=True)
    return n.cuda()


def test_full_scan_descriptor_vs_oracle(sd, net):
    """One full 65 536-point scan: bit-exact pyramid and tables, descriptor within 1e-4 of the torch-CPU oracle."""
    from lcrnet_b200 import data as gdata
    raw = synth.make_scan(5, 7356)
    assert raw.shape == (65536, 3)
    d = gdata.scans_collate_fn_stack_mode([raw], 4, 0.3, 1.275, LIMITS, pre_voxel=0.3, int32=False)
    p0, l0 = on.grid_subsample(raw, np.array([len(raw)], dtype=np.int64), 0.3)
    ref = mo.precompute_pyramid(p0, l0, limits=LIMITS)
    for key in ('points', 'neighbors', 'subsampling'):
        for a, b in zip(d[key], ref[key]):
            assert torch.equal(a.cpu(), b), key
    with torch.no_grad():
        want = mo.global_descriptor(sd, ref).numpy()
    got = net(d)['anc_global'].cpu().numpy()
    rel = float(np.linalg.norm(got - want) / np.linalg.norm(want))
    print('full-scan descriptor relative L2 error vs oracle: %.2e' % rel)
    assert rel < 1e-4


def test_bench_batch_properties(net):
    """The bench configuration at reduced batch (8 full scans, 2 streams): unit-norm descriptors, every scan equal
    to its own single-scan forward (scans are independent units), tables sorted and inside the radius."""
    from lcrnet_b200 import data as gdata
    from lcrnet_b200 import pipeline
    scans = [synth.make_scan(i // 2, 7351 + i) for i in range(8)]
    pts = torch.from_numpy(np.concatenate(scans, 0)).pin_memory()
    pipe = pipeline.DescriptorPipeline(net, LIMITS, n_streams=2)
    desc = pipe(pts, [len(s) for s in scans])
    torch.cuda.synchronize()
    pipe.close()
    assert desc.shape == (8, 256) and torch.isfinite(desc).all()
    assert float((desc.norm(dim=1) - 1).abs().max()) < 1e-5
    for i in (0, 5):
        one = gdata.scans_collate_fn_stack_mode([scans[i]], 4, 0.3, 1.275, LIMITS, pre_voxel=0.3)
        single = net(one)['anc_global'][0]
        assert float((single - desc[i]).norm()) < 1e-5
    d = gdata.scans_collate_fn_stack_mode(scans[:4], 4, 0.3, 1.275, LIMITS, pre_voxel=0.3, int32=True)
    r = 1.275
    for level in range(4):
        p, nb = d['points'][level], d['neighbors'][level].long()
        n = p.shape[0]
        valid = nb < n
        q = p[:, None, :]
        s = torch.cat([p, torch.full((1, 3), 1e6, device=p.device)])[nb]
        d2 = ((q - s) ** 2).sum(-1)
        assert bool((d2[valid] < r * r * (1 + 1e-6)).all())                       # inside the radius
        assert bool((nb[:, 0] == torch.arange(n, device=p.device)).all())         # nearest neighbour = the point itself
        dd = torch.where(valid, d2, torch.full_like(d2, float('inf')))
        assert bool((dd[:, 1:] >= dd[:, :-1] - 1e-6).all())                       # ascending distance, pads last
        lens = d['lengths'][level].tolist()                                        # neighbours never cross clouds
        cloud = torch.repeat_interleave(torch.arange(len(lens), device=p.device), torch.tensor(lens, device=p.device))
        cl = torch.cat([cloud, torch.tensor([-1], device=p.device)])[nb]
        assert bool(((cl == cloud[:, None]) | ~valid).all())
        r *= 2


def test_database_topk_at_config3_size():
    """4000 x 4000 brute-force top-25 (configs[3]): self match first, ascending distances, a sample of rows exact
    against the numpy restatement of the faiss search."""
    from lcrnet_b200 import ops
    rng = np.random.default_rng(12)
    db = rng.standard_normal((4000, 256)).astype(np.float32)
    db /= np.linalg.norm(db, axis=1, keepdims=True)
    dbg = torch.from_numpy(db).cuda()
    d2, idx = ops.l2_topk(dbg, dbg, 25)
    assert bool((idx[:, 0] == torch.arange(4000, device=idx.device)).all())
    assert float(d2[:, 0].abs().max()) < 1e-5
    assert bool((d2[:, 1:] >= d2[:, :-1]).all())
    rows = rng.choice(4000, 64, replace=False)
    ref_d, ref_i = mo.l2_topk(db[rows], db, 25)
    assert np.allclose(d2[rows].cpu().numpy(), ref_d, rtol=1e-5, atol=1e-6)
    gap = np.abs(np.diff(ref_d, axis=1)) > 1e-5
    ok = np.ones_like(ref_i, dtype=bool)
    ok[:, 1:] &= gap
    ok[:, :-1] &= gap
    assert np.array_equal(idx[rows].cpu().numpy()[ok], ref_i[ok])
