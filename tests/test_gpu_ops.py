"""GPU parity: the CUDA operators (through the C ABI / utils.ext drop-in) against the C oracle
and the committed reference fixtures.  Bit-exact: integer / index work and fp32 centroids."""
import os

import numpy as np
import pytest
import torch

from oracle import native as on
from util import GOLDEN, canonical_rows, random_clouds

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(GOLDEN, 'ops_golden.npz'))


def _gpu_subsample(pts, lens, voxel, order='reference'):
    from lcrnet_b200 import ext
    p, l = ext.grid_subsampling(torch.from_numpy(pts).cuda(), torch.from_numpy(lens).cuda(), voxel, order=order)
    return p.cpu().numpy(), l.cpu().numpy()


def _gpu_radius(q, s, ql, sl, r, **kw):
    from lcrnet_b200 import ext
    t = ext.radius_neighbors(torch.from_numpy(q).cuda(), torch.from_numpy(s).cuda(), torch.from_numpy(ql).cuda(),
                             torch.from_numpy(sl).cuda(), r, **kw)
    return t.cpu().numpy()


@pytest.mark.parametrize('name', ['a', 'b', 'c'])
def test_golden_fixtures(name):
    pts, lens = random_clouds(int(G[name + '_seed']), list(G[name + '_sizes']), extent=8.0, z_extent=2.0)
    voxel, radius = float(G[name + '_voxel']), float(G[name + '_radius'])
    s_pts, s_lens = _gpu_subsample(pts, lens, voxel)
    assert np.array_equal(s_lens, G[name + '_s_lengths'])
    assert np.array_equal(s_pts.view(np.uint32), G[name + '_s_points'].view(np.uint32))
    for q, ql, key in ((pts, lens, '_neighbors'), (s_pts, s_lens, '_subsampling')):
        ref = G[name + key].astype(np.int64)
        got = _gpu_radius(q, pts, ql, lens, radius)
        assert got.shape == ref.shape and got.dtype == np.int64
        d_ref = on.neighbor_d2(q, pts, ref)
        assert np.array_equal(canonical_rows(ref, d_ref)[0], got)


@pytest.mark.parametrize('seed,sizes,voxel', [(1, [3000], 0.6), (2, [2000, 1, 2500], 1.2), (3, [20000], 0.3),
                                               (4, [12, 13, 14000], 2.4), (5, [50, 0, 20], 5.0),
                                               (6, [7] * 40, 1.0)])
def test_subsample_vs_oracle_bit_exact(seed, sizes, voxel):
    pts, lens = random_clouds(seed, sizes, extent=40.0)
    a, al = on.grid_subsample(pts, lens, voxel)
    b, bl = _gpu_subsample(pts, lens, voxel)
    assert np.array_equal(al, bl)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    # first-seen order mode: same set of centroids, bit-identical values
    c, cl = _gpu_subsample(pts, lens, voxel, order='first_seen')
    assert np.array_equal(al, cl)
    o = 0
    for n in al:
        x, y = a[o:o + n].view(np.uint32), c[o:o + n].view(np.uint32)
        assert np.array_equal(x[np.lexsort(x.T)], y[np.lexsort(y.T)])
        o += n


def test_subsample_synthetic_scan_pyramid():
    from lcrnet_b200 import synth
    pts = synth.make_scan(0)
    lens = np.array([len(pts)], dtype=np.int64)
    voxel = 0.3
    for _ in range(4):
        a, al = on.grid_subsample(pts, lens, voxel)
        b, bl = _gpu_subsample(pts, lens, voxel)
        assert np.array_equal(al, bl) and np.array_equal(a.view(np.uint32), b.view(np.uint32))
        pts, lens, voxel = a, al, voxel * 2


@pytest.mark.parametrize('seed,sizes,radius', [(1, [1500], 2.0), (2, [800, 900], 3.0), (3, [1, 5, 0, 700], 4.0),
                                               (4, [3000, 2500], 1.0)])
def test_radius_vs_oracle_exact(seed, sizes, radius):
    pts, lens = random_clouds(seed, sizes)
    ref, counts, mc = on.radius_neighbors(pts, pts, lens, lens, radius, return_counts=True)
    got = _gpu_radius(pts, pts, lens, lens, radius)
    assert got.shape == ref.shape
    assert np.array_equal(got, ref)  # same tie-break (ascending support index): exact equality
    # fused limit cut == python slice of the full table; int32 variant identical
    cut = _gpu_radius(pts, pts, lens, lens, radius, limit=9)
    assert np.array_equal(cut, ref[:, :9])
    cut32 = _gpu_radius(pts, pts, lens, lens, radius, limit=9, int32=True)
    assert cut32.dtype == np.int32 and np.array_equal(cut32.astype(np.int64), ref[:, :9])
    wide = _gpu_radius(pts, pts, lens, lens, radius, limit=mc + 50)
    assert wide.shape[1] == mc  # narrower than the limit, like the reference table


def test_radius_query_support_differ_and_spill_path():
    # dense cloud: > 512 neighbours per query exercises the CTA spill kernel
    rng = np.random.default_rng(9)
    s = rng.uniform(-2, 2, (6000, 3)).astype(np.float32)
    q = rng.uniform(-2, 2, (300, 3)).astype(np.float32)
    ql, sl = np.array([300], dtype=np.int64), np.array([6000], dtype=np.int64)
    ref, counts, mc = on.radius_neighbors(q, s, ql, sl, 1.5, return_counts=True)
    assert mc > 512
    got = _gpu_radius(q, s, ql, sl, 1.5)
    assert np.array_equal(got, ref)


def test_pyramid_tables_on_synthetic_scan():
    from lcrnet_b200 import synth
    raw = synth.make_scan(1)
    p0, l0 = on.grid_subsample(raw, np.array([len(raw)], dtype=np.int64), 0.3)
    p1, l1 = on.grid_subsample(p0, l0, 0.6)
    for q, ql, s, sl, r in ((p0, l0, p0, l0, 1.275), (p1, l1, p0, l0, 1.275), (p0, l0, p1, l1, 2.55)):
        ref = on.radius_neighbors(q, s, ql, sl, r, limit=60)
        got = _gpu_radius(q, s, ql, sl, r, limit=60)
        assert np.array_equal(got, ref)


def test_cpu_tensors_round_trip_like_reference_module():
    from lcrnet_b200 import ext
    pts, lens = random_clouds(7, [500])
    p, l = ext.grid_subsampling(torch.from_numpy(pts), torch.from_numpy(lens), 1.0)
    assert not p.is_cuda and not l.is_cuda
    a, al = on.grid_subsample(pts, lens, 1.0)
    assert np.array_equal(p.numpy().view(np.uint32), a.view(np.uint32))
    t = ext.radius_neighbors(torch.from_numpy(pts), torch.from_numpy(pts), torch.from_numpy(lens),
                             torch.from_numpy(lens), 2.0)
    assert not t.is_cuda and np.array_equal(t.numpy(), on.radius_neighbors(pts, pts, lens, lens, 2.0))


def test_cell_grouped_self_search_forced():
    """The cell-grouped self-table kernel (radius.cu query_self_kernel) only takes over above 200 k supports; forced
    on (LCR_RADIUS_CELLS=2, read once per process -> subprocess) it must reproduce the oracle bit for bit on the small
    clouds of this file too, including rows that overflow the warp buffer (spill path) and the counting pass."""
    import subprocess
    import sys
    code = r'''
import sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
import numpy as np, torch
from lcrnet_b200 import ext
from oracle import native as on
from util import random_clouds
for seed, sizes, radius in ((3, [4000, 2500, 1], 1.2), (4, [9000], 3.0), (5, [300, 5000, 700], 0.7)):
    pts, lens = random_clouds(seed, sizes)
    p, l = torch.from_numpy(pts).cuda(), torch.from_numpy(lens).cuda()
    for limit in (0, 40):
        got = ext.radius_neighbors(p, p, l, l, radius, limit=limit).cpu().numpy()
        want = on.radius_neighbors(pts, pts, lens, lens, radius, limit if limit else None)
        assert got.shape == want.shape and np.array_equal(got, want), (seed, limit, got.shape, want.shape)
print('CELLS-OK')
''' % (REPO, os.path.join(REPO, 'tests'))
    env = dict(os.environ, LCR_RADIUS_CELLS='2')
    r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and 'CELLS-OK' in r.stdout, r.stderr[-3000:]


@pytest.mark.gpu
def test_nearest_only_search_is_column_zero():
    """Nearest-only mode (the up-sampling tables of the registration pipeline): one column, equal to column 0 of the
    full sorted table incl. the rows without any support inside the radius, for int32 and int64 tables, with and
    without a reused support grid."""
    import torch
    from lcrnet_b200 import ext, ops
    g = torch.Generator().manual_seed(11)
    sl = torch.tensor([700, 1300], dtype=torch.int64)
    ql = torch.tensor([2500, 1800], dtype=torch.int64)
    s = (torch.rand(int(sl.sum()), 3, generator=g) * 20).cuda()
    q = (torch.rand(int(ql.sum()), 3, generator=g) * 24 - 2).cuda()          # some queries outside the support box
    q[:50] = s[:50]                                                            # exact hits (d2 = 0)
    ql_d, sl_d = ql.cuda(), sl.cuda()
    for int32 in (True, False):
        full = ops.radius_search(q, s, ql_d, sl_d, 1.7, 40, int32=int32)
        near = ops.radius_search(q, s, ql_d, sl_d, 1.7, 40, int32=int32, nearest=True)
        assert near.shape == (q.shape[0], 1) and near.dtype == full.dtype
        assert torch.equal(near[:, 0], full[:, 0])
        assert int((near[:, 0] == s.shape[0]).sum()) > 0                       # pads exist
        grid = ext.SupportGrid(s, sl_d, 1.7, q.shape[0])
        a = ops.radius_search(q, s, ql_d, sl_d, 1.7, 40, int32=int32, grid=grid)
        b = ops.radius_search(q, s, ql_d, sl_d, 1.7, 40, int32=int32, grid=grid, nearest=True)
        assert torch.equal(a, full) and torch.equal(b, near)
