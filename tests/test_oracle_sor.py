"""oracle/sor.py (restatement of PCL's StatisticalOutlierRemoval as the reference configures it,
src/rosFuncs.cpp:20-24) against an independent exact kNN (scipy.spatial.cKDTree, float64).  PCL itself is
not available here: this pins the kNN part and the statistics, not PCL's bits."""
import numpy as np
import pytest
from scipy.spatial import cKDTree

from oracle import sor


def _cloud(n, seed):
    rng = np.random.default_rng(seed)
    # a road-scene-like cloud: ground plane, two walls, scattered outliers; z forward up to 60 m
    g = np.c_[rng.uniform(-10, 10, n // 2), rng.normal(1.65, 0.02, n // 2), rng.uniform(4, 60, n // 2)]
    w = np.c_[rng.choice([-8.0, 8.0], n // 3) + rng.normal(0, 0.05, n // 3), rng.uniform(-3, 1.6, n // 3),
              rng.uniform(4, 60, n // 3)]
    o = np.c_[rng.uniform(-30, 30, n - n // 2 - n // 3), rng.uniform(-10, 5, n - n // 2 - n // 3),
              rng.uniform(-20, 120, n - n // 2 - n // 3)]
    return np.concatenate([g, w, o]).astype(np.float32)


@pytest.mark.parametrize("n,k", [(3000, 200), (1200, 50)])
def test_mean_knn_distances_match_kdtree(n, k):
    p = _cloud(n, n)
    d, valid = sor.mean_knn_distances(p, k)
    assert valid == n
    dd, _ = cKDTree(p.astype(np.float64)).query(p.astype(np.float64), k + 1)
    ref = dd[:, 1:].mean(1)
    assert np.abs(d - ref).max() <= 2e-6 * ref.max()


def test_sor_cloud_semantics():
    p = _cloud(3000, 1)
    p[5] = (0, 0, -600)            # -z > 500: never enters the cloud (src/rosFuncs.cpp:12)
    keep, dist, thr = sor.sor_cloud(p, 200, 0.01, return_all=True)
    assert 5 not in keep and dist[5] == -1
    inside = dist >= 0
    assert np.array_equal(keep, np.nonzero(inside & ~(dist > thr))[0])
    assert 0.3 * inside.sum() < len(keep) < inside.sum()            # mul = 0.01: the sparse tail goes
    assert np.all(np.diff(keep) > 0)
    # fewer points than meanK + 1: no distance is valid, the threshold is NaN and everything is kept
    q = p[:150]
    assert np.array_equal(sor.sor_cloud(q, 200, 0.01), np.nonzero(~(-q[:, 2] > 500))[0])
