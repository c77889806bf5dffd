"""Candidate-vertex sequences for the GSkel replay tests (CPU and GPU): noisy skeleton points of a branching structure seen
through a window that sweeps upwards.  With sigma = 1.3 m and 0.9 m per tick the reference's state machine goes through
every event: spawn, Kalman fusion, approval, the > 50 % new abort, distance and joint-joint merges, leaf pruning,
Laplacian smoothing, isolated (NaN) vertices, branch extraction."""
import numpy as np


def structure_ticks(n_ticks=70, seed=0, sigma=1.3, step=0.9, keep=0.8, spacing=1.1):
    rng = np.random.default_rng(seed)
    segs = [((0, 0, 0), (0, 0, 40)), ((0, 0, 40), (18, 0, 62)), ((0, 0, 40), (-16, 6, 60)), ((0, 0, 22), (12, 0, 26)), ((0, 0, 30), (-3, 9, 33))]
    pts = []
    for a, b in segs:
        a = np.array(a, float); b = np.array(b, float)
        L = np.linalg.norm(b - a)
        for s in np.arange(0, L, spacing):
            pts.append(a + (b - a) * s / L)
    pts = np.array(pts)
    out = []
    for k in range(n_ticks):
        zc = 8 + step * k
        m = (pts[:, 2] > zc - 30) & (pts[:, 2] < zc + 6)
        sel = pts[m]
        sel = sel[rng.random(len(sel)) < keep]
        c = sel + rng.normal(0, sigma, sel.shape)
        out.append(c[rng.permutation(len(c))].astype(np.float32))
    return out


def weird_clouds(fx):
    """degenerate geometries shared by the GPU parity tests and the CPU oracle-vs-pyref tests"""
    rng = np.random.default_rng(77)
    out = {}
    base = fx.cylinder(8000, seed=12)
    # isolated outliers far from everything (kNN slow path: ring search beyond R = 2, brute force)
    outl = rng.uniform(-40, 40, (12, 3)).astype(np.float32)
    out["outliers"] = np.concatenate([base, outl])
    # planar patch (normals all parallel -> planar branch of the position step, rosa.cpp:334)
    pl = np.stack([rng.uniform(5, 15, 6000), rng.uniform(-5, 5, 6000), 0.01 * rng.normal(size=6000)], 1).astype(np.float32)
    out["plane"] = pl
    # thin line: 1-D structure, the leaf loop oscillates / runs out of passes
    t = rng.uniform(0, 30, 5000)
    out["line"] = np.stack([10 + 0.01 * rng.normal(size=5000), 0.01 * rng.normal(size=5000), t - 15], 1).astype(np.float32)
    # two far-apart clusters (huge bbox, tiny occupied volume)
    a = rng.normal(size=(3000, 3)).astype(np.float32) * 0.5 + np.array([10, -30, -20], np.float32)
    b = rng.normal(size=(3000, 3)).astype(np.float32) * 0.5 + np.array([12, 30, 25], np.float32)
    out["two_clusters"] = np.concatenate([a, b])
    # heavy duplication: 400 distinct points, each repeated 10 times (distance ties everywhere)
    d = fx.cylinder(400, seed=5)
    out["duplicates"] = np.repeat(d, 10, axis=0)
    # coarse lattice (many exactly equal distances)
    g = np.stack(np.meshgrid(np.arange(10, 14, 0.25), np.arange(-3, 3, 0.25), np.arange(-4, 4, 0.25), indexing="ij"), -1).reshape(-1, 3)
    out["lattice"] = g.astype(np.float32)
    return out
