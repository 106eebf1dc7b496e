"""Independent float64 model of the reference ground segmentation (numpy only, no oracle code).

This is the only pin available at the Eigen boundary (reference src/segmentation.cpp:62-102): Eigen 3.4 is
neither vendored nor installed, so neither the restated oracle (oracle/oracle.cpp: float32, sequential sums,
restated JacobiSVD) nor the device (double-precision moments, the same restated JacobiSVD in float32) can be
compared bit for bit with the reference's own arithmetic. What CAN be stated is how far both are from the
mathematically exact plane of the same point set: this model follows src/segmentation.cpp:104-309 step by
step, keeps the integer/ordering logic exact (stable x order, float32 z comparisons, the ascending
SEQUENTIAL float32 sum of the lowest representatives, :189-197) and evaluates every plane fit and every
point-to-plane distance in float64 with numpy.linalg.eigh.

Sign of the normal: Eigen's JacobiSVD never flips the columns of V and every rotation keeps c > 0, so a
ground-like covariance (z is the thin axis) gives normal.z > 0; the model fixes normal.z > 0 and
the tests assert that the oracle and the device agree with that on every fitted plane.
"""
from __future__ import annotations

import numpy as np

UNKNOWN, GROUND, OBSTACLE = 0, 1, 2


def segment_f64(pts, sensor_height_m=1.73, orthogonal_distance_threshold=0.3, initial_seed_threshold=0.6,
                number_of_iterations=3, number_of_planar_partitions=2, number_of_lower_point_representatives=5000):
    """Returns dict(labels uint32[N], planes float64[P, iterations, 4] (a, b, c, d; NaN where no fit ran),
    n_ground int64[P, iterations], margin float64[N] = signed distance of every classified point to the decision
    surface of its partition's LAST plane (dist - thr; NaN for unclassified points))."""
    pts = np.ascontiguousarray(pts, np.float32)
    n = pts.shape[0]
    P, iters = int(number_of_planar_partitions), int(number_of_iterations)
    labels = np.zeros(n, np.uint32)
    planes = np.full((P, iters, 4), np.nan)
    n_ground = np.zeros((P, iters), np.int64)
    margin = np.full(n, np.nan)
    if n == 0:
        return dict(labels=labels, planes=planes, n_ground=n_ground, margin=margin)
    order = np.argsort(pts[:, 0], kind="stable")  # src/segmentation.cpp:119 with ties by index
    per = n // P
    thr32 = np.float32(orthogonal_distance_threshold)
    for s in range(P):
        idx = order[s * per:min((s + 1) * per, n)]
        m = idx.size
        if m < 3:  # :225-229, points stay UNKNOWN
            continue
        z = pts[idx, 2]
        zo = np.argsort(z, kind="stable")  # :165
        zs = z[zo]
        z_min = np.float32(-1.5) * np.float32(sensor_height_m)
        above = np.nonzero(zs > z_min)[0]
        cut = int(above[0]) if above.size else 0  # :171-182
        zo, zs = zo[cut:], zs[cut:]
        ground = np.empty(0, np.int64)
        if zs.size:
            k = min(zs.size, int(number_of_lower_point_representatives))
            mean = np.cumsum(zs[:k], dtype=np.float32)[-1] / np.float32(k)  # sequential float32 sum, :189-197
            z_max = np.float32(mean + np.float32(initial_seed_threshold))
            above = np.nonzero(zs > z_max)[0]
            ground = np.sort(zo[:int(above[0]) if above.size else 0])  # :199-216 (no point above -> no seeds)
        X = pts[idx, :3].astype(np.float64)
        failed = False
        dist = thr = None
        for it in range(iters):
            if ground.size < 3:  # :251-259
                failed = True
                break
            G = X[ground]
            c = G.mean(axis=0)
            C = G - c
            cov = C.T @ C / (ground.size - 1)
            w, v = np.linalg.eigh(cov)
            normal = v[:, 0]
            if normal[2] < 0:
                normal = -normal
            d = float(normal @ c)
            planes[s, it] = (*normal, d)
            n_ground[s, it] = ground.size
            dist = X @ normal - d  # :290-291
            thr = float(thr32) * float(np.linalg.norm(normal))  # :293
            ground = np.nonzero(dist < thr)[0]  # :297-307
        if failed:
            labels[idx] = OBSTACLE
            continue
        lab = np.full(m, OBSTACLE, np.uint32)
        lab[ground] = GROUND
        labels[idx] = lab
        margin[idx] = dist - thr
    return dict(labels=labels, planes=planes, n_ground=n_ground, margin=margin)


def plane_deviation(plane, plane_f64):
    """(max |normal component difference|, |d difference|) of a float plane against the float64 model's; both are
    compared after normalising the float normal (the decision rule scales the threshold by its norm)."""
    p = np.asarray(plane, np.float64)
    q = np.asarray(plane_f64, np.float64)
    nrm = np.linalg.norm(p[:3])
    return float(np.max(np.abs(p[:3] / nrm - q[:3]))), float(abs(p[3] / nrm - q[3]))
