"""CPU: the RANSAC oracle's pieces -- the sampler restatement is the product's (splitmix hash), Kabsch recovers a
rigid motion, the estimator finds it under outliers."""
import numpy as np

from oracle import ransac_oracle as ro


def test_sampler_is_deterministic_and_in_range():
    xs = [ro.draw(3, h, k, 1000) for h in range(50) for k in range(3)]
    assert min(xs) >= 0 and max(xs) < 1000 and len(set(xs)) > 100
    assert xs == [ro.draw(3, h, k, 1000) for h in range(50) for k in range(3)]
    assert ro.draw(0, 0, 0, 7919) == (ro._mix64(0) >> 11) % 7919               # sample = hash(seed ^ f(h, k)) % n


def test_oracle_recovers_motion():
    rng = np.random.default_rng(1)
    src = rng.uniform(-10, 10, (400, 3)).astype(np.float32)
    a = 0.4
    R = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
    t = np.array([1.0, -2.0, 0.3])
    ref = (src @ R.T + t).astype(np.float32)
    ref[::3] += rng.uniform(-5, 5, ref[::3].shape).astype(np.float32)
    T, h, c = ro.ransac(src, ref, 0.05, 3, 300, seed=1)
    assert c >= 250 and np.abs(T[:3, :3] - R).max() < 1e-3 and np.abs(T[:3, 3] - t).max() < 1e-2
