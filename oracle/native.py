"""ctypes loaders for the CPU checker libraries (TEST INFRASTRUCTURE ONLY).

* ``liblcr_oracle.so`` -- oracle/lcr_oracle.c, the plain-C restatement of
  ``grid_subsampling`` / ``radius_neighbors`` (reference: utils/extensions/cpu/**).
* ``_ref/libref_ext.so`` / ``_ref/libref_legacy.so`` -- the unmodified reference sources,
  compiled where they lie by oracle/Makefile (only present when built in the container
  that has /root/reference; they travel to the GPU box as prebuilt files).

Nothing under ``lcr-net_b200/`` imports this module.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_i64p = ctypes.POINTER(ctypes.c_int64)
_i32p = ctypes.POINTER(ctypes.c_int32)
_f32p = ctypes.POINTER(ctypes.c_float)


def build(quiet=True):
    """Compile the checker libraries (make is incremental)."""
    subprocess.run(['make', '-C', _HERE], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _load(path):
    if not os.path.exists(path):
        return None
    return ctypes.CDLL(path)


_oracle = None
_ref = None
_legacy = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        p = os.path.join(_HERE, '_build', 'liblcr_oracle.so')
        if not os.path.exists(p):
            build()
        _oracle = ctypes.CDLL(p)
        _oracle.lcr_oracle_grid_subsample.restype = ctypes.c_int64
        _oracle.lcr_oracle_grid_subsample.argtypes = [_f32p, _i64p, ctypes.c_int, ctypes.c_float, _f32p, _i64p]
        _oracle.lcr_oracle_radius_neighbors.restype = ctypes.c_int64
        _oracle.lcr_oracle_radius_neighbors.argtypes = [_f32p, _f32p, _i64p, _i64p, ctypes.c_int, ctypes.c_float,
                                                        ctypes.c_int64, _i64p, _i32p]
        _oracle.lcr_oracle_neighbor_d2.restype = None
        _oracle.lcr_oracle_neighbor_d2.argtypes = [_f32p, _f32p, ctypes.c_int64, ctypes.c_int64, _i64p,
                                                   ctypes.c_int64, _f32p]
    return _oracle


def ref_lib():
    """The compiled reference (or None if it was not built / shipped)."""
    global _ref
    if _ref is None:
        _ref = _load(os.path.join(_HERE, '_ref', 'libref_ext.so'))
        if _ref is not None:
            _ref.ref_grid_subsampling.restype = ctypes.c_int64
            _ref.ref_grid_subsampling.argtypes = [_f32p, ctypes.c_int64, _i64p, ctypes.c_int, ctypes.c_float,
                                                  _f32p, _i64p]
            _ref.ref_radius_neighbors.restype = ctypes.c_int64
            _ref.ref_radius_neighbors.argtypes = [_f32p, ctypes.c_int64, _f32p, ctypes.c_int64, _i64p, _i64p,
                                                  ctypes.c_int, ctypes.c_float, _i64p]
    return _ref


def legacy_lib():
    global _legacy
    if _legacy is None:
        _legacy = _load(os.path.join(_HERE, '_ref', 'libref_legacy.so'))
        if _legacy is not None:
            _legacy.ref_batch_ordered_neighbors.restype = ctypes.c_int64
            _legacy.ref_batch_ordered_neighbors.argtypes = [_f32p, ctypes.c_int64, _f32p, ctypes.c_int64, _i64p,
                                                            _i64p, ctypes.c_int, ctypes.c_float, _i32p]
            _legacy.ref_legacy_grid_subsampling.restype = ctypes.c_int64
            _legacy.ref_legacy_grid_subsampling.argtypes = [_f32p, ctypes.c_int64, ctypes.c_float, _f32p]
    return _legacy


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _ptr(a, t):
    return a.ctypes.data_as(t)


# --------------------------------------------------------------------------- oracle (C port)
def grid_subsample(points, lengths, voxel):
    """Restatement of utils.ext.grid_subsampling (pybind.cpp:14-18): (s_points, s_lengths)."""
    points, lengths = _f32(points).reshape(-1, 3), _i64(lengths)
    out = np.empty_like(points)
    out_len = np.empty_like(lengths)
    m = oracle_lib().lcr_oracle_grid_subsample(_ptr(points, _f32p), _ptr(lengths, _i64p), len(lengths),
                                               float(voxel), _ptr(out, _f32p), _ptr(out_len, _i64p))
    if m < 0:
        raise RuntimeError('oracle grid_subsample failed: %d' % m)
    return out[:m].copy(), out_len


def radius_neighbors(q, s, q_len, s_len, radius, limit=None, return_counts=False):
    """Restatement of utils.ext.radius_neighbors (pybind.cpp:9-13) + the python slice
    ``[:, :limit]`` of ops/radius_search.py:25-26.  Ties: ascending support index."""
    q, s = _f32(q).reshape(-1, 3), _f32(s).reshape(-1, 3)
    q_len, s_len = _i64(q_len), _i64(s_len)
    lib = oracle_lib()
    counts = np.zeros(len(q), dtype=np.int32)
    mc = lib.lcr_oracle_radius_neighbors(_ptr(q, _f32p), _ptr(s, _f32p), _ptr(q_len, _i64p), _ptr(s_len, _i64p),
                                         len(q_len), float(radius), 0, None, _ptr(counts, _i32p))
    width = mc if (limit is None or limit <= 0) else min(mc, limit)
    out = np.empty((len(q), width), dtype=np.int64)
    lib.lcr_oracle_radius_neighbors(_ptr(q, _f32p), _ptr(s, _f32p), _ptr(q_len, _i64p), _ptr(s_len, _i64p),
                                    len(q_len), float(radius), width, _ptr(out, _i64p), None)
    if return_counts:
        return out, counts, mc
    return out


def neighbor_d2(q, s, idx):
    q, s, idx = _f32(q).reshape(-1, 3), _f32(s).reshape(-1, 3), _i64(idx)
    out = np.empty(idx.shape, dtype=np.float32)
    oracle_lib().lcr_oracle_neighbor_d2(_ptr(q, _f32p), _ptr(s, _f32p), idx.shape[0], s.shape[0],
                                        _ptr(idx, _i64p), idx.shape[1], _ptr(out, _f32p))
    return out


# --------------------------------------------------------------------------- compiled reference
def ref_grid_subsample(points, lengths, voxel):
    lib = ref_lib()
    if lib is None:
        raise RuntimeError('oracle/_ref/libref_ext.so not available')
    points, lengths = _f32(points).reshape(-1, 3), _i64(lengths)
    out = np.empty_like(points)
    out_len = np.empty_like(lengths)
    m = lib.ref_grid_subsampling(_ptr(points, _f32p), len(points), _ptr(lengths, _i64p), len(lengths),
                                 float(voxel), _ptr(out, _f32p), _ptr(out_len, _i64p))
    return out[:m].copy(), out_len


def ref_radius_neighbors(q, s, q_len, s_len, radius, limit=None):
    lib = ref_lib()
    if lib is None:
        raise RuntimeError('oracle/_ref/libref_ext.so not available')
    q, s = _f32(q).reshape(-1, 3), _f32(s).reshape(-1, 3)
    q_len, s_len = _i64(q_len), _i64(s_len)
    args = (_ptr(q, _f32p), len(q), _ptr(s, _f32p), len(s), _ptr(q_len, _i64p), _ptr(s_len, _i64p),
            len(q_len), float(radius))
    mc = lib.ref_radius_neighbors(*args, None)
    out = np.empty((len(q), mc), dtype=np.int64)
    lib.ref_radius_neighbors(*args, _ptr(out, _i64p))
    if limit is not None and limit > 0:
        out = np.ascontiguousarray(out[:, :limit])
    return out


def ref_batch_ordered_neighbors(q, s, q_len, s_len, radius):
    lib = legacy_lib()
    if lib is None:
        raise RuntimeError('oracle/_ref/libref_legacy.so not available')
    q, s = _f32(q).reshape(-1, 3), _f32(s).reshape(-1, 3)
    q_len, s_len = _i64(q_len), _i64(s_len)
    args = (_ptr(q, _f32p), len(q), _ptr(s, _f32p), len(s), _ptr(q_len, _i64p), _ptr(s_len, _i64p),
            len(q_len), float(radius))
    mc = lib.ref_batch_ordered_neighbors(*args, None)
    out = np.empty((len(q), mc), dtype=np.int32)
    lib.ref_batch_ordered_neighbors(*args, _ptr(out, _i32p))
    return out.astype(np.int64)
