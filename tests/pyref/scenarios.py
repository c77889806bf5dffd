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
