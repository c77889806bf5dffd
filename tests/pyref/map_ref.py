"""CPU restatement of the accumulated-map semantics (include/osep_skel.h: osep_map_*; csrc/map.cu), TEST INFRASTRUCTURE.
A voxel of side `voxel` keeps the first point that ever fell into it (lowest frame, then lowest index within the frame);
the kept points are listed in exactly that order.  Points are first moved to the map frame with tf2_sensor_msgs'
float32 arithmetic (gskel_node.cpp:137-138): R = Quaternionf(w, x, y, z).toRotationMatrix() in float32, p' = (R p) + t with
the three products of a row summed as a + (b + c)."""
import numpy as np

F = np.float32


def tf_points(xyz, translation, rotation_xyzw):
    p = np.asarray(xyz, F)[:, :3]
    if translation is None or rotation_xyzw is None:
        return p.copy()
    x, y, z, w = [F(v) for v in rotation_xyzw]
    tx, ty, tz = F(2) * x, F(2) * y, F(2) * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    R = np.array([[F(1) - (tyy + tzz), txy - twz, txz + twy],
                  [txy + twz, F(1) - (txx + tzz), tyz - twx],
                  [txz - twy, tyz + twx, F(1) - (txx + tyy)]], F)
    T = np.array(translation, np.float64).astype(F)
    out = np.empty_like(p)
    for r in range(3):
        a, b, c = R[r, 0] * p[:, 0], R[r, 1] * p[:, 1], R[r, 2] * p[:, 2]
        out[:, r] = (a + (b + c)) + T[r]
    return out


class MapRef:
    def __init__(self, voxel):
        self.inv = F(1.0) / F(voxel)
        self.seen = set()
        self.points = []

    def insert(self, xyz, translation=None, rotation_xyzw=None):
        g = tf_points(xyz, translation, rotation_xyzw)
        k = np.floor(g * self.inv)
        ok = np.isfinite(g).all(1) & (np.abs(k) < 1048575.0).all(1)
        fresh = 0
        ki = k.astype(np.int64)
        for i in np.nonzero(ok)[0]:
            key = (int(ki[i, 0]), int(ki[i, 1]), int(ki[i, 2]))
            if key not in self.seen:
                self.seen.add(key); self.points.append(g[i]); fresh += 1
        return fresh

    def array(self):
        return np.array(self.points, F).reshape(-1, 3)
