"""Shared helpers for the parity tests."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
REF_PRESENT = os.path.isdir('/root/reference/utils/extensions')


def random_clouds(seed, sizes, extent=20.0, z_extent=3.0):
    """Stacked random clouds with duplicated points and lattice points (forces exact distance
    ties and multi-point voxels)."""
    rng = np.random.default_rng(seed)
    clouds = []
    for n in sizes:
        p = rng.uniform(-1.0, 1.0, (n, 3)) * np.array([extent, extent, z_extent])
        k = n // 8
        if k:
            p[:k] = np.round(p[:k] * 2.0) / 2.0       # lattice points -> exact ties
            p[k:2 * k] = p[:k]                           # exact duplicates
        clouds.append(p.astype(np.float32))
    pts = np.concatenate(clouds, 0) if clouds else np.zeros((0, 3), np.float32)
    return pts, np.array(sizes, dtype=np.int64)


def canonical_rows(idx, d2):
    """Sort every row by (d2, idx): two tables that agree up to the order inside exact-distance
    tie classes have identical canonical forms."""
    order = np.lexsort((idx, d2), axis=1)
    return np.take_along_axis(idx, order, 1), np.take_along_axis(d2, order, 1)
