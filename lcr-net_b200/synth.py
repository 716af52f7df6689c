"""Synthetic KITTI-shaped LiDAR scans: an HDL-64E-like ray cast,
64 beams x 1024 azimuth steps = 65 536 returns per scan, float32 xyz.

The world of a ``scene_seed`` ("a place") is a ground plane, a smooth closed street-canyon wall
and axis-aligned boxes; ``seed`` fixes the range noise.  A pair is two scans of the same scene
from two sensor poses, so the ground-truth relative pose is known.
Deterministic: numpy ``default_rng`` (PCG64) streams only.
"""
import numpy as np

N_BEAMS = 64
N_AZIMUTH = 1024
SENSOR_HEIGHT = 1.73
MAX_RANGE = 80.0


def _scene(scene_seed, n_boxes):
    rng = np.random.default_rng([int(scene_seed), 0x5CE7E])
    amp = rng.uniform(2.0, 9.0, 4)
    phase = rng.uniform(0.0, 2 * np.pi, 4)
    centres = rng.uniform(-70.0, 70.0, (n_boxes, 2))
    half = rng.uniform(1.0, 5.0, (n_boxes, 2))
    height = rng.uniform(1.0, 4.0, n_boxes)
    # keep a 10 m clearing around the origin so no sensor pose of a pair starts inside a box
    d = np.linalg.norm(centres, axis=1)
    need = 10.0 + np.linalg.norm(half, axis=1)
    centres = centres * np.maximum(1.0, need / np.maximum(d, 1e-6))[:, None]
    return amp, phase, centres, half, height


def make_scan(scene_seed=0, seed=7351, pose=(0.0, 0.0, 0.0), n_boxes=100, noise=0.02):
    """One scan in the SENSOR frame.  pose = (x, y, yaw) of the sensor in the scene frame.
    Returns float32 [65536, 3]."""
    amp, phase, centres, half, height = _scene(scene_seed, n_boxes)
    px, py, yaw = pose
    elev = np.deg2rad(np.linspace(2.0, -24.8, N_BEAMS))
    azim = np.linspace(0.0, 2 * np.pi, N_AZIMUTH, endpoint=False)
    el, az = np.meshgrid(elev, azim, indexing='ij')
    el, az = el.ravel(), az.ravel()
    # ray directions in the scene frame
    dx, dy, dz = np.cos(el) * np.cos(az + yaw), np.cos(el) * np.sin(az + yaw), np.sin(el)
    t = np.full(el.shape, MAX_RANGE)
    # ground plane z = 0, sensor at height SENSOR_HEIGHT
    down = dz < -1e-6
    tg = np.where(down, -SENSOR_HEIGHT / np.where(down, dz, -1.0), np.inf)
    t = np.minimum(t, tg)
    # street-canyon wall: radial distance R(theta) from the scene origin; solved per ray by a few
    # fixed-point steps on the horizontal range (smooth, slowly varying R)
    hd = np.sqrt(dx * dx + dy * dy)
    s = np.full(el.shape, 60.0)
    for _ in range(6):
        wx, wy = px + dx / hd * s, py + dy / hd * s
        th = np.arctan2(wy, wx)
        R = 60.0 + sum(amp[k] * np.sin((k + 1) * th + phase[k]) for k in range(4))
        R = np.clip(R, 35.0, 79.0)
        rho = np.sqrt(wx * wx + wy * wy)
        s = np.maximum(s + (R - rho), 0.5)
    tw = s / hd
    zw = SENSOR_HEIGHT + dz * tw
    tw = np.where((zw >= 0.0) & (zw <= 12.0), tw, np.inf)
    t = np.minimum(t, tw)
    # boxes (slab test), standing on the ground
    ox, oy, oz = px, py, SENSOR_HEIGHT
    with np.errstate(divide='ignore', invalid='ignore'):
        for c, h, hz in zip(centres, half, height):
            lo = np.array([c[0] - h[0], c[1] - h[1], 0.0])
            hi = np.array([c[0] + h[0], c[1] + h[1], hz])
            t1x, t2x = (lo[0] - ox) / dx, (hi[0] - ox) / dx
            t1y, t2y = (lo[1] - oy) / dy, (hi[1] - oy) / dy
            t1z, t2z = (lo[2] - oz) / dz, (hi[2] - oz) / dz
            tmin = np.maximum(np.maximum(np.minimum(t1x, t2x), np.minimum(t1y, t2y)), np.minimum(t1z, t2z))
            tmax = np.minimum(np.minimum(np.maximum(t1x, t2x), np.maximum(t1y, t2y)), np.maximum(t1z, t2z))
            hit = (tmax >= np.maximum(tmin, 0.0)) & (tmin > 0.3)
            t = np.where(hit, np.minimum(t, tmin), t)
    rng = np.random.default_rng([int(scene_seed), int(seed), 0xA015E])
    t = np.minimum(t + rng.normal(0.0, noise, t.shape), MAX_RANGE)
    # back to the sensor frame (undo yaw)
    sx, sy = np.cos(el) * np.cos(az) * t, np.cos(el) * np.sin(az) * t
    sz = dz * t
    return np.stack([sx, sy, sz], axis=1).astype(np.float32)


def make_pair(scene_seed=0, seed=7351, max_shift=4.0):
    """(ref_scan, src_scan, T_src_to_ref[4,4]) of one scene from two poses."""
    rng = np.random.default_rng([int(scene_seed), int(seed), 0x9A12])
    shift = rng.uniform(-max_shift, max_shift, 2)
    yaw = rng.uniform(-np.pi, np.pi)
    ref = make_scan(scene_seed, seed, (0.0, 0.0, 0.0))
    src = make_scan(scene_seed, seed + 1, (shift[0], shift[1], yaw))
    T = np.eye(4)
    T[:2, :2] = [[np.cos(yaw), -np.sin(yaw)], [np.sin(yaw), np.cos(yaw)]]
    T[:2, 3] = shift
    return ref, src, T.astype(np.float32)
