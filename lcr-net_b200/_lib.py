"""ctypes binding of ``liblcr_b200.so`` (the C ABI of include/lcr_b200.h).

The library is the only compute path of this package: if it cannot be loaded, or CUDA is not
available when an operator is called, a RuntimeError is raised (never a CPU fallback).
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'liblcr_b200.so')

c_i32, c_i64, c_f32, c_sz, c_vp = ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_size_t, ctypes.c_void_p

# name -> (restype, argtypes); must list every symbol include/lcr_b200.h declares
SIGNATURES = {
    'lcr_last_error': (ctypes.c_char_p, []),
    'lcr_abi_version': (c_i32, []),
    'lcr_launch_count': (c_i64, []),
    'lcr_profile_begin': (None, []),
    'lcr_profile_end': (c_i32, []),
    'lcr_profile_get': (c_i32, [c_i32, ctypes.c_char_p, c_i32, ctypes.POINTER(ctypes.c_double),
                                ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]),
    'lcr_grid_subsample_ws_bytes': (c_sz, [c_i64, c_i32]),
    'lcr_grid_subsample': (c_i32, [c_vp, c_i64, c_vp, c_i32, c_f32, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
    'lcr_radius_neighbors_ws_bytes': (c_sz, [c_i64, c_i64, c_i32]),
    'lcr_radius_neighbors': (c_i32, [c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_i32, c_f32, c_i32, c_vp, c_i32, c_vp,
                                     c_vp, c_vp, c_vp, c_sz, c_vp]),
    'lcr_radius_neighbors_ex': (c_i32, [c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_i32, c_f32, c_i32, c_vp, c_i32, c_vp,
                                        c_vp, c_vp, c_vp, c_sz, c_i32, c_vp]),
    'lcr_kpconv_ws_bytes': (c_sz, [c_i64, c_i32]),
    'lcr_kpconv_ws_bytes2': (c_sz, [c_i64, c_i64, c_i32]),
    'lcr_kpconv': (c_i32, [c_vp, c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_i32, c_i32, c_vp, c_vp, c_f32, c_vp, c_vp,
                           c_vp, c_vp, c_i32, c_i32, c_vp, c_vp, c_sz, c_vp]),
    'lcr_kpconv_gn': (c_i32, [c_vp, c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_i32, c_i32, c_vp, c_vp, c_f32, c_vp, c_vp,
                              c_vp, c_vp, c_i32, c_i32, c_vp, c_vp, c_sz, c_vp, c_vp, c_i32, c_vp]),
    'lcr_set_gather_mode': (None, [c_i32]),
    'lcr_row_flags': (c_i32, [c_vp, c_i64, c_i32, c_vp, c_vp]),
    'lcr_linear': (c_i32, [c_vp, c_i64, c_i32, c_vp, c_i32, c_vp, c_vp, c_vp]),
    'lcr_group_norm_ws_bytes': (c_sz, [c_i64, c_i32, c_i32]),
    'lcr_group_norm_stats': (c_i32, [c_vp, c_i64, c_i32, c_i32, c_vp, c_i32, c_i64, c_f32, c_vp, c_vp, c_sz, c_vp]),
    'lcr_gn_blocks_ws_bytes': (c_sz, [c_i64, c_i32]),
    'lcr_group_norm_finalize_blocks': (c_i32, [c_vp, c_i64, c_i32, c_i32, c_vp, c_i32, c_f32, c_vp, c_vp]),
    'lcr_group_norm_apply': (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_i32, c_vp,
                                     c_i32, c_i32, c_f32, c_vp, c_vp, c_vp]),
    'lcr_maxpool': (c_i32, [c_vp, c_i64, c_vp, c_i32, c_i32, c_i64, c_i32, c_vp, c_vp]),
    'lcr_netvlad_ws_bytes': (c_sz, [c_i64, c_i32]),
    'lcr_netvlad': (c_i32, [c_vp, c_i64, c_vp, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_sz,
                            c_vp]),
    'lcr_linear_ex': (c_i32, [c_vp, c_i64, c_i32, c_i32, c_vp, c_i32, c_vp, c_vp, c_i32, c_vp, c_i32, c_vp]),
    'lcr_linear_tc': (c_i32, [c_vp, c_i64, c_i32, c_i32, c_vp, c_vp, c_i32, c_i32, c_vp, c_vp, c_i32, c_vp, c_i32,
                              c_vp]),
    'lcr_linear_tc_gn': (c_i32, [c_vp, c_i64, c_i32, c_i32, c_vp, c_vp, c_i32, c_i32, c_vp, c_vp, c_i32, c_vp, c_vp,
                                 c_i32, c_vp]),
    'lcr_tf32_split': (c_i32, [c_vp, c_i64, c_vp, c_vp, c_vp]),
    'lcr_layer_norm': (c_i32, [c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_f32, c_i32, c_vp, c_vp]),
    'lcr_rope': (c_i32, [c_vp, c_i32, c_vp, c_i64, c_vp]),
    'lcr_attention': (c_i32, [c_vp, c_i32, c_vp, c_i32, c_vp, c_i32, c_vp, c_vp, c_i32, c_i64, c_i32, c_i32, c_vp,
                              c_i32, ctypes.c_double, c_vp]),
    'lcr_attention_tc': (c_i32, [c_vp, c_i32, c_vp, c_i32, c_vp, c_i32, c_vp, c_vp, c_i32, c_i64, c_i32, c_i32, c_vp,
                                 c_i32, ctypes.c_double, c_vp]),
    'lcr_attention_tma': (c_i32, [c_vp, c_i32, c_i64, c_vp, c_i32, c_vp, c_i32, c_i64, c_vp, c_vp, c_i32, c_i64, c_i32, c_i32,
                                  c_vp, c_i32, ctypes.c_double, c_vp]),
    'lcr_vote_shift': (c_i32, [c_vp, c_vp, c_i32, c_f32, c_i64, c_vp, c_vp]),
    'lcr_nms_greedy': (c_i32, [c_vp, c_vp, c_i32, c_i64, c_f32, c_vp, c_vp, c_vp, c_vp]),
    'lcr_neighbor_mean': (c_i32, [c_vp, c_i64, c_vp, c_i32, c_i32, c_i64, c_vp, c_vp]),
    'lcr_upsample_concat': (c_i32, [c_vp, c_i64, c_i32, c_vp, c_i32, c_vp, c_i32, c_i64, c_vp, c_vp]),
    'lcr_point_to_node_ws_bytes': (c_sz, [c_i64, c_i64]),
    'lcr_point_to_node': (c_i32, [c_vp, c_i64, c_vp, c_i64, c_i32, c_vp, c_vp, c_vp, c_i32, c_vp, c_vp, c_vp, c_sz,
                                  c_vp]),
    'lcr_sinkhorn': (c_i32, [c_vp, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_i32, c_vp, c_vp]),
    'lcr_sinkhorn_stats': (c_i32, [ctypes.POINTER(c_i64), c_i32]),
    'lcr_set_sinkhorn_cluster': (None, [c_i32]),
    'lcr_set_sinkhorn_early_exit': (None, [c_i32]),
    'lcr_coarse_matching_ws_bytes': (c_sz, [c_i32, c_i32]),
    'lcr_coarse_matching': (c_i32, [c_vp, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
    'lcr_patch_scores': (c_i32, [c_vp, c_i64, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_i32, c_i32, c_i32, c_vp, c_vp]),
    'lcr_fine_correspondences_ws_bytes': (c_sz, [c_i32]),
    'lcr_fine_correspondences': (c_i32, [c_vp, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp,
                                         c_vp, c_sz, c_vp]),
    'lcr_corr_points': (c_i32, [c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'lcr_lgr_ws_bytes': (c_sz, [c_i32, c_i64]),
    'lcr_local_global_registration': (c_i32, [c_vp, c_vp, c_vp, c_vp, c_i32, c_i64, c_f32, c_i32, c_i32, c_vp, c_vp,
                                              c_sz, c_vp]),
    'lcr_point_to_node_batched': (c_i32, [c_vp, c_i64, c_vp, c_vp, c_i64, c_vp, c_i32, c_i64, c_i64, c_i32, c_vp, c_vp, c_vp,
                                          c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
    'lcr_node_scores': (c_i32, [c_vp, c_i32, c_vp, c_vp, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp]),
    'lcr_coarse_matching_batched': (c_i32, [c_vp, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
    'lcr_gather_coarse': (c_i32, [c_vp, c_vp, c_vp, c_i32, c_vp, c_vp, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'lcr_lgr_batched_ws_bytes': (c_sz, [c_i32, c_i32, c_i64]),
    'lcr_lgr_batched': (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_vp, c_vp, c_i32, c_i64, c_f32, c_i32, c_i32, c_vp,
                                c_vp, c_sz, c_vp]),
    'lcr_ransac_ws_bytes': (c_sz, [c_i32]),
    'lcr_ransac_correspondences': (c_i32, [c_vp, c_vp, c_i64, c_f32, c_i32, c_i32, ctypes.c_uint64, c_vp, c_vp, c_vp, c_sz, c_vp]),
    'lcr_l2_topk': (c_i32, [c_vp, c_i64, c_vp, c_i64, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp]),
}

_lib = None


def lib():
    """The loaded library (loads on first use; raises if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError('%s is missing: run `python -c "import __graft_entry__ as g; g.build()"` '
                               '(there is no CPU fallback)' % LIB_PATH)
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc):
    if rc != 0:
        raise RuntimeError('lcr_b200 error %d: %s' % (rc, lib().lcr_last_error().decode()))


def require_cuda(*tensors):
    if not torch.cuda.is_available():
        raise RuntimeError('lcrnet_b200 needs a CUDA device (sm_100a); there is no CPU fallback')
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError('expected a CUDA tensor')


def ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


_raw_stream = getattr(torch._C, '_cuda_getCurrentRawStream', None)


def stream_ptr(device=None):
    """cudaStream_t of torch's current stream on ``device`` (every operator is queued there).  The raw accessor is
    ~10x cheaper than building a torch.cuda.Stream object: this is called once per kernel launch."""
    if _raw_stream is not None:
        idx = getattr(device, 'index', device)
        if not isinstance(idx, int):
            idx = torch.cuda.current_device()
        return ctypes.c_void_p(_raw_stream(idx))
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class _Workspace:
    """Grow-only scratch buffers, one per (device, stream, slot): work queued on different
    streams (pipeline.py) never shares scratch memory."""

    def __init__(self):
        self._buf = {}

    def get(self, nbytes, device, slot=0):
        idx = device.index if device.index is not None else torch.cuda.current_device()
        key = (idx, _raw_stream(idx) if _raw_stream is not None else torch.cuda.current_stream(device).cuda_stream,
               slot)
        b = self._buf.get(key)
        if b is None or b.numel() < nbytes:
            b = torch.empty(max(int(nbytes * 1.25), 1 << 20), dtype=torch.uint8, device=device)
            self._buf[key] = b
        return b


workspace = _Workspace()
