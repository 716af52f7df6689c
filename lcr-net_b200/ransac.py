"""RANSAC registration from correspondences on the GPU -- the reference's alternative estimator
(``utils/utils/open3d.py:145-173``, used by ``experiments/registration/eval.py:176``), same name and arguments.

The reference delegates to open3d's ``registration_ransac_based_on_correspondence``; open3d is third party and
not in the reference tree (parity unpinned at that boundary).  ``liblcr_b200.so`` implements the published
algorithm with all hypotheses in parallel (csrc/ransac.cu); ``seed`` makes every hypothesis reproducible."""
import numpy as np
import torch

from . import _lib


def ransac_from_correspondences(src_corr_points, ref_corr_points, distance_threshold=0.05, ransac_n=3,
                                num_iterations=10000, seed=0):
    """Device API: src / ref float32 CUDA [n, 3] (row t is one correspondence) ->
    (T float32 [4, 4] on the device mapping src -> ref, best hypothesis index, its inlier count)."""
    _lib.require_cuda(src_corr_points, ref_corr_points)
    src = src_corr_points.to(torch.float32).contiguous()
    ref = ref_corr_points.to(torch.float32).contiguous()
    assert src.shape == ref.shape and src.dim() == 2 and src.shape[1] == 3 and src.shape[0] >= 1
    L = _lib.lib()
    T = torch.empty((4, 4), dtype=torch.float32, device=src.device)
    best = torch.empty(2, dtype=torch.int32, device=src.device)
    ws = _lib.workspace.get(L.lcr_ransac_ws_bytes(int(num_iterations)), src.device, slot=7)
    _lib.check(L.lcr_ransac_correspondences(_lib.ptr(src), _lib.ptr(ref), src.shape[0], float(distance_threshold),
                                            int(ransac_n), int(num_iterations), int(seed) & (2 ** 64 - 1), _lib.ptr(T),
                                            _lib.ptr(best), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(src.device)))
    return T, best


def registration_with_ransac_from_correspondences(src_points, ref_points, correspondences=None,
                                                  distance_threshold=0.05, ransac_n=3, num_iterations=10000, seed=0):
    """utils/utils/open3d.py:145-173, same arguments: numpy points [N, 3] (+ optional index pairs [M, 2] into
    src / ref; default: row i <-> row i) -> 4 x 4 float64 numpy transform from src to ref."""
    src = np.asarray(src_points, dtype=np.float32)
    ref = np.asarray(ref_points, dtype=np.float32)
    if correspondences is not None:
        c = np.asarray(correspondences, dtype=np.int64)
        src, ref = src[c[:, 0]], ref[c[:, 1]]
    T, _ = ransac_from_correspondences(torch.from_numpy(np.ascontiguousarray(src)).cuda(),
                                       torch.from_numpy(np.ascontiguousarray(ref)).cuda(), distance_threshold, ransac_n,
                                       num_iterations, seed)
    return T.cpu().numpy().astype(np.float64)
