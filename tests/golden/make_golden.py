"""Generates the committed fixtures under tests/golden/ from the UNMODIFIED reference, in the
build container (needs /root/reference):

  ops_golden.npz   -- utils/extensions C++ (compiled by oracle/Makefile into oracle/_ref):
                      grid_subsampling + radius_neighbors on small seeded clouds.
  model_golden.npz -- the reference Python model (imported through tests/golden/ref_import.py)
                      on a small synthetic scan with seeded weights (see make_model_golden()).

Run:  python tests/golden/make_golden.py [ops|model|all]
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'tests'))

from oracle import native as on  # noqa: E402
from util import random_clouds  # noqa: E402


def make_ops_golden():
    on.build()
    assert on.ref_lib() is not None, 'oracle/_ref/libref_ext.so missing (needs /root/reference)'
    out = {}
    cases = [('a', 11, [700, 500], 0.6, 1.5), ('b', 12, [1, 300], 1.2, 2.5), ('c', 13, [1500], 0.3, 0.9)]
    for name, seed, sizes, voxel, radius in cases:
        pts, lens = random_clouds(seed, sizes, extent=8.0, z_extent=2.0)
        s_pts, s_lens = on.ref_grid_subsample(pts, lens, voxel)
        nbr = on.ref_radius_neighbors(pts, pts, lens, lens, radius)
        sub = on.ref_radius_neighbors(s_pts, pts, s_lens, lens, radius)
        out.update({name + '_seed': seed, name + '_sizes': np.array(sizes), name + '_voxel': np.float32(voxel),
                    name + '_radius': np.float32(radius), name + '_s_points': s_pts, name + '_s_lengths': s_lens,
                    name + '_neighbors': nbr.astype(np.int32), name + '_subsampling': sub.astype(np.int32)})
    np.savez_compressed(os.path.join(HERE, 'ops_golden.npz'), **out)
    print('wrote ops_golden.npz', os.path.getsize(os.path.join(HERE, 'ops_golden.npz')))


if __name__ == '__main__':
    what = sys.argv[1] if len(sys.argv) > 1 else 'all'
    if what in ('ops', 'all'):
        make_ops_golden()
    if what in ('model', 'all'):
        from make_model_golden import make_model_golden
        make_model_golden()
