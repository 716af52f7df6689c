"""CPU restatement (numpy) of the RANSAC-from-correspondences estimator -- TEST INFRASTRUCTURE ONLY.

Reference call site: utils/utils/open3d.py:145-173 (open3d registration_ransac_based_on_correspondence with
TransformationEstimationPointToPoint(False)); open3d is third party and absent: PARITY UNPINNED at that boundary.
This restates the published algorithm with the product's counter-based sampler so both evaluate the same
hypotheses: sample ransac_n correspondences, Kabsch on the sample, count |T src - ref| < threshold, best count
wins (ties: lower squared residual sum of the inliers, then lower hypothesis index)."""
import numpy as np

M64 = (1 << 64) - 1


def _mix64(x):
    x ^= x >> 33
    x = (x * 0xff51afd7ed558ccd) & M64
    x ^= x >> 33
    x = (x * 0xc4ceb9fe1a85ec53) & M64
    x ^= x >> 33
    return x


def draw(seed, h, k, n):
    x = _mix64((seed ^ ((((h << 8) | k) * 0x9E3779B97F4A7C15) & M64)) & M64)
    return (x >> 11) % n


def kabsch(src, ref):
    sc, rc = src.mean(0), ref.mean(0)
    H = (src - sc).T @ (ref - rc)
    U, _, Vt = np.linalg.svd(H)
    d = np.sign(np.linalg.det(Vt.T @ U.T))
    R = Vt.T @ np.diag([1.0, 1.0, d]) @ U.T
    return R, rc - R @ sc


def ransac(src, ref, distance_threshold=0.05, ransac_n=3, num_iterations=10000, seed=0, hypotheses=None):
    """-> (T [4,4] float64, best hypothesis, inlier count).  ``hypotheses``: evaluate only these indices."""
    src64, ref64 = np.asarray(src, np.float64), np.asarray(ref, np.float64)
    n = len(src64)
    best = (-1, np.inf, -1, None)
    for h in (range(num_iterations) if hypotheses is None else hypotheses):
        idx = [draw(seed, h, k, n) for k in range(ransac_n)]
        R, t = kabsch(src64[idx], ref64[idx])
        R32, t32 = R.astype(np.float32), t.astype(np.float32)
        d2 = (((np.asarray(src, np.float32) @ R32.T + t32) - np.asarray(ref, np.float32)) ** 2).sum(1)
        inl = d2 < np.float32(distance_threshold) ** 2
        c, e = int(inl.sum()), float(d2[inl].sum())
        if c > best[0] or (c == best[0] and e < best[1]):
            T = np.eye(4)
            T[:3, :3], T[:3, 3] = R, t
            best = (c, e, h, T)
    return best[3], best[2], best[0]
