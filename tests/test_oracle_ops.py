"""CPU: the C oracle (oracle/lcr_oracle.c) against (1) the committed fixtures generated from the
compiled reference and (2), when /root/reference is present, the compiled reference itself."""
import os

import numpy as np
import pytest

from oracle import native as on
from util import GOLDEN, REF_PRESENT, canonical_rows, random_clouds

G = np.load(os.path.join(GOLDEN, 'ops_golden.npz'))
CASES = ['a', 'b', 'c']


def _case(name):
    pts, lens = random_clouds(int(G[name + '_seed']), list(G[name + '_sizes']), extent=8.0, z_extent=2.0)
    return pts, lens, float(G[name + '_voxel']), float(G[name + '_radius'])


@pytest.mark.parametrize('name', CASES)
def test_subsample_matches_golden_bit_exact(name):
    pts, lens, voxel, _ = _case(name)
    s_pts, s_lens = on.grid_subsample(pts, lens, voxel)
    assert np.array_equal(s_lens, G[name + '_s_lengths'])
    # bit-exact values AND the reference's unordered_map iteration order
    assert np.array_equal(s_pts.view(np.uint32), G[name + '_s_points'].view(np.uint32))


@pytest.mark.parametrize('name', CASES)
def test_radius_matches_golden_modulo_ties(name):
    pts, lens, voxel, radius = _case(name)
    s_pts, s_lens = G[name + '_s_points'], G[name + '_s_lengths']
    for q, ql, key in ((pts, lens, '_neighbors'), (s_pts, s_lens, '_subsampling')):
        ref = G[name + key].astype(np.int64)
        got = on.radius_neighbors(q, pts, ql, lens, radius)
        assert got.shape == ref.shape
        d_ref, d_got = on.neighbor_d2(q, pts, ref), on.neighbor_d2(q, pts, got)
        assert np.array_equal(d_ref, d_got)            # identical distance rows
        ci, _ = canonical_rows(ref, d_ref)
        assert np.array_equal(ci, got)                 # identical up to order inside tie classes
        assert (got[d_got == np.inf] == len(pts)).all()  # pad value = total support count


def test_radius_limit_cut_and_counts():
    pts, lens = random_clouds(5, [400, 300])
    full, counts, mc = on.radius_neighbors(pts, pts, lens, lens, 4.0, return_counts=True)
    assert full.shape[1] == mc == counts.max()
    cut = on.radius_neighbors(pts, pts, lens, lens, 4.0, limit=7)
    assert np.array_equal(cut, full[:, :7])
    assert ((full < len(pts)).sum(1) == counts).all()
    # neighbours never cross batch elements
    assert (full[:400][full[:400] < len(pts)] < 400).all()
    assert (full[400:][full[400:] < len(pts)] >= 400).all()


def test_empty_cloud_in_batch():
    pts, lens = random_clouds(6, [50, 0, 20])
    s_pts, s_lens = on.grid_subsample(pts, lens, 5.0)
    assert s_lens[1] == 0 and s_lens.sum() == len(s_pts)
    idx = on.radius_neighbors(pts, pts, lens, lens, 3.0)
    assert idx.shape[0] == 70


@pytest.mark.skipif(not REF_PRESENT, reason='reference tree not present')
@pytest.mark.parametrize('seed,sizes,voxel', [(1, [3000], 0.6), (2, [2000, 1, 2500], 1.2), (3, [20000], 0.3),
                                               (4, [12, 13, 14000], 2.4)])
def test_subsample_vs_compiled_reference(seed, sizes, voxel):
    pts, lens = random_clouds(seed, sizes, extent=40.0)
    a, al = on.grid_subsample(pts, lens, voxel)
    b, bl = on.ref_grid_subsample(pts, lens, voxel)
    assert np.array_equal(al, bl)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.skipif(not REF_PRESENT, reason='reference tree not present')
@pytest.mark.parametrize('seed,sizes,radius', [(1, [1500], 2.0), (2, [800, 900], 3.0)])
def test_radius_vs_compiled_reference_and_legacy(seed, sizes, radius):
    pts, lens = random_clouds(seed, sizes)
    got = on.radius_neighbors(pts, pts, lens, lens, radius)
    ref = on.ref_radius_neighbors(pts, pts, lens, lens, radius)
    assert got.shape == ref.shape
    d_ref = on.neighbor_d2(pts, pts, ref)
    ci, _ = canonical_rows(ref, d_ref)
    assert np.array_equal(ci, got)
    if on.legacy_lib() is not None:
        # batch_ordered_neighbors: the reference's own deterministic (ascending index) tie order.
        # It pads with ns and indexes supports globally as well.
        leg = on.ref_batch_ordered_neighbors(pts, pts, lens, lens, radius)
        assert leg.shape == got.shape
        d_leg = on.neighbor_d2(pts, pts, leg)
        assert np.array_equal(d_leg, on.neighbor_d2(pts, pts, got))
        assert np.array_equal(canonical_rows(leg, d_leg)[0], got)


@pytest.mark.skipif(not REF_PRESENT, reason='reference tree not present')
def test_demo_scan_pyramid_vs_compiled_reference():
    scan = np.load('/root/reference/demo/data_demo/003854.npy')[:, :3].astype(np.float32)
    pts, lens = np.ascontiguousarray(scan), np.array([len(scan)], dtype=np.int64)
    voxel = 0.6
    for _ in range(3):
        a, al = on.grid_subsample(pts, lens, voxel)
        b, bl = on.ref_grid_subsample(pts, lens, voxel)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)) and np.array_equal(al, bl)
        pts, lens, voxel = a, al, voxel * 2
