"""Synthetic structure-like LiDAR clouds (SURVEY.md Appendix B / BASELINE.md section 3).

The reference ships no data (no .pcd/.bag, no tests); these generators define the
workloads BASELINE.json's configs name.  All outputs are float32 (N, 3), sensor at the
origin, isotropic Gaussian noise sigma = 0.02 m, numpy.random.default_rng(seed); the draw
order below is part of the specification (it fixes every fixture bit-for-bit).
"""
import numpy as np

SIGMA = 0.02


def _cyl(rng, n, r, h, cx, z0, sigma=SIGMA):
    th = rng.uniform(np.pi / 2, 3 * np.pi / 2, n)     # LiDAR-visible half (faces -x, towards the sensor)
    z = rng.uniform(z0, z0 + h, n)
    P = np.stack([cx + r * np.cos(th), r * np.sin(th), z], axis=1)
    return P + rng.normal(0.0, sigma, (n, 3))


def _blade(rng, n, L, hub, ang, sigma=SIGMA):
    t = rng.uniform(0, 1, n)
    th = rng.uniform(np.pi / 2, 3 * np.pi / 2, n)
    a = 1.5 * (1 - t) + 0.2
    b = 0.4 * (1 - t) + 0.1
    d = np.array([0.0, np.cos(ang), np.sin(ang)])
    e = np.array([0.0, -np.sin(ang), np.cos(ang)])
    x = np.array([1.0, 0.0, 0.0])
    P = np.asarray(hub)[None, :] + (t * L)[:, None] * d + (a * np.sin(th))[:, None] * e + (b * np.cos(th))[:, None] * x
    return P + rng.normal(0.0, sigma, (n, 3))


def _turbine_local(rng, n, cx):
    n_t = n // 2
    n_b = (n - n_t) // 3
    n_last = n - n_t - 2 * n_b
    hub = (cx, 0.0, 20.0)
    parts = [_cyl(rng, n_t, 2.0, 45.0, cx, -25.0)]
    for k in range(3):
        parts.append(_blade(rng, n_b if k < 2 else n_last, 28.0, hub, np.pi / 2 + k * 2 * np.pi / 3))
    return np.concatenate(parts, axis=0)


def cylinder(n=20000, seed=0, r=2.0, h=50.0, cx=15.0, z0=-25.0):
    """C1 'cyl20k': visible half of a vertical cylinder."""
    rng = np.random.default_rng(seed)
    return _cyl(rng, n, r, h, cx, z0).astype(np.float32)


def turbine(n=100000, seed=0, cx=20.0, yaw=0.0):
    """C2 'turbine100k': tower (50 %) + 3 tapered blades 120 deg apart from hub (cx, 0, 20)."""
    rng = np.random.default_rng(seed)
    P = _turbine_local(rng, n, cx)
    if yaw != 0.0:
        c, s = np.cos(yaw), np.sin(yaw)
        R = np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])
        P = P @ R.T
    return P.astype(np.float32)


def batch_frame(b, n=100000):
    """C4 frame b: C2 with seed b, random sensor yaw and +-5 m stand-off jitter."""
    rng = np.random.default_rng(1_000_003 + b)
    yaw = float(rng.uniform(0.0, 2 * np.pi))
    cx = 20.0 + float(rng.uniform(-5.0, 5.0))
    return turbine(n=n, seed=b, cx=cx, yaw=yaw)


def stream_frame(f, n=50000, n_frames=1000):
    """C3 frame f: turbine at the odom origin seen from a sensor on a helix (radius 25 m, one
    revolution per 250 frames, z from -10 to +15 m).  Returns (points in the sensor frame
    float32 (n,3), R (3,3), t (3,)) with p_odom = R @ p_sensor + t."""
    rng = np.random.default_rng(f)
    phi = 2 * np.pi * f / 250.0
    t = np.array([25.0 * np.cos(phi), 25.0 * np.sin(phi), -10.0 + 25.0 * f / max(1, n_frames - 1)])
    xa = np.array([-np.cos(phi), -np.sin(phi), 0.0])    # sensor x-axis points at the tower axis
    za = np.array([0.0, 0.0, 1.0])
    ya = np.cross(za, xa)
    R = np.stack([xa, ya, za], axis=1)
    Pl = _turbine_local(rng, n, 0.0)                     # local -x is the visible side
    c, s = np.cos(phi + np.pi), np.sin(phi + np.pi)      # rotate local -x onto the tower->sensor direction
    Rz = np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])
    P_odom = Pl @ Rz.T
    P_sensor = (P_odom - t[None, :]) @ R
    return P_sensor.astype(np.float32), R, t


def to_xyz4(P):
    """pcl::PointXYZ layout: 16 B per point {x, y, z, 1.0f}."""
    out = np.ones((P.shape[0], 4), dtype=np.float32)
    out[:, :3] = P
    return out


def quat_xyzw_from_matrix(R):
    """Unit quaternion (x, y, z, w) of a rotation matrix -- the rotation field of the geometry_msgs/TransformStamped a
    TF lookup returns for the analytic sensor pose of stream_frame()."""
    R = np.asarray(R, dtype=np.float64)
    tr = R[0, 0] + R[1, 1] + R[2, 2]
    if tr > 0:
        s = 2.0 * np.sqrt(tr + 1.0); w = 0.25 * s
        x = (R[2, 1] - R[1, 2]) / s; y = (R[0, 2] - R[2, 0]) / s; z = (R[1, 0] - R[0, 1]) / s
    elif R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]:
        s = 2.0 * np.sqrt(1.0 + R[0, 0] - R[1, 1] - R[2, 2]); x = 0.25 * s
        w = (R[2, 1] - R[1, 2]) / s; y = (R[0, 1] + R[1, 0]) / s; z = (R[0, 2] + R[2, 0]) / s
    elif R[1, 1] > R[2, 2]:
        s = 2.0 * np.sqrt(1.0 + R[1, 1] - R[0, 0] - R[2, 2]); y = 0.25 * s
        w = (R[0, 2] - R[2, 0]) / s; x = (R[0, 1] + R[1, 0]) / s; z = (R[1, 2] + R[2, 1]) / s
    else:
        s = 2.0 * np.sqrt(1.0 + R[2, 2] - R[0, 0] - R[1, 1]); z = 0.25 * s
        w = (R[1, 0] - R[0, 1]) / s; x = (R[0, 2] + R[2, 0]) / s; y = (R[1, 2] + R[2, 1]) / s
    return np.array([x, y, z, w], dtype=np.float64)
