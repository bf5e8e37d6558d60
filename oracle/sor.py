"""Restatement of visualSLAM::SORcloud (reference src/rosFuncs.cpp:9-39).  TEST INFRASTRUCTURE ONLY.

The reference drops points with -z > 500 and runs pcl::StatisticalOutlierRemoval (meanK = 200,
stddevMulThresh = 0.01) on the rest.  PCL is a third-party dependency that the reference does not vendor
(`find_package(PCL 1.8 REQUIRED)`, CMakeLists.txt) and that is not installed in the build image, so the
algorithm is restated here from PCL's published implementation
(filters/include/pcl/filters/impl/statistical_outlier_removal.hpp, PCL 1.8 .. 1.12, `applyFilterIndices`):

  first pass   for every finite point: nearestKSearch(point, meanK + 1) on a FLANN kd-tree (exact search,
               squared L2 distances accumulated in float, L2_Simple<float>), skip neighbour 0 (the query
               itself), dist_sum (double) += sqrt(nn_dists[k]) (float square root), distance =
               (float)(dist_sum / meanK).  A point whose meanK + 1 neighbours cannot be found keeps
               distance 0 and is not counted.
  statistics   sum += d; sq_sum += d * d (float product) over all points in double;
               mean = sum / valid; variance = (sq_sum - sum * sum / valid) / (valid - 1);
               threshold = mean + stddevMul * sqrt(variance)
  second pass  keep points with distance <= threshold (a NaN threshold keeps everything)

PARITY UNPINNED: there is no PCL here to run, and the reference has no test vectors for this step.  The
restatement is cross-checked against scipy.spatial.cKDTree (an independent exact kNN in float64) in
tests/test_oracle_sor.py: identical neighbour sets up to float ties, mean distances equal to float precision.
"""
import numpy as np


def mean_knn_distances(pts, mean_k=200, chunk=512):
    """Mean distance of every point to its mean_k nearest neighbours, PCL arithmetic (float squared
    distances ((dx*dx)+dy*dy)+dz*dz, float sqrt, double sum in ascending order, float result)."""
    pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 3)
    n = len(pts)
    out = np.zeros(n, np.float32)
    if n <= mean_k:
        return out, 0
    x, y, z = pts[:, 0], pts[:, 1], pts[:, 2]
    for s in range(0, n, chunk):
        q = pts[s:s + chunk]
        dx = q[:, 0:1] - x[None, :]
        dy = q[:, 1:2] - y[None, :]
        dz = q[:, 2:3] - z[None, :]
        d2 = (dx * dx + dy * dy) + dz * dz                     # float32 throughout, this association
        part = np.partition(d2, mean_k, axis=1)[:, :mean_k + 1]
        part.sort(axis=1)
        r = np.sqrt(part[:, 1:])                               # float32 sqrt; column 0 is the query itself
        acc = np.zeros(len(q), np.float64)
        for k in range(r.shape[1]):                            # PCL's order: ascending distance
            acc += r[:, k].astype(np.float64)
        out[s:s + chunk] = (acc / float(mean_k)).astype(np.float32)
    return out, n


def sor_cloud(ref3d, mean_k=200, stddev_mul=0.01, return_all=False):
    """visualSLAM::SORcloud: returns the input indices of the kept points (input order)."""
    p = np.ascontiguousarray(ref3d, np.float32).reshape(-1, 3)
    cloud = np.nonzero(~(-1 * p[:, 2] > 500))[0]               # src/rosFuncs.cpp:12
    pc = p[cloud]
    finite = np.isfinite(pc).all(1)
    dist = np.zeros(len(cloud), np.float32)
    d, valid = mean_knn_distances(pc[finite], mean_k)
    dist[finite] = d
    s = 0.0
    sq = 0.0
    for v in dist:                                             # sequential, double, float product
        s += float(v)
        sq += float(np.float32(v) * np.float32(v))
    with np.errstate(all="ignore"):
        mean = np.float64(s) / np.float64(valid)
        var = (np.float64(sq) - np.float64(s) * np.float64(s) / np.float64(valid)) / (np.float64(valid) - 1.0)
        thr = mean + np.float64(stddev_mul) * np.sqrt(var)
    keep = cloud[~(dist.astype(np.float64) > thr)]
    if return_all:
        full = np.full(len(p), -1.0, np.float32)
        full[cloud] = dist
        return keep.astype(np.int32), full, float(thr)
    return keep.astype(np.int32)
