"""Accumulated map on the device (include/osep_skel.h: osep_map_*) -- the input of the tiled configuration.
Frames are transformed into the map frame and inserted into a voxel hash; a voxel keeps the first point that fell into it,
and the kept points form an append-only dense device array in (frame, index) order."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import OsepError


def _dbl(a, n):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1)
    assert a.size == n
    return a


class SkelMap:
    def __init__(self, device=0, capacity_points=1 << 22, max_frame_points=1 << 20, voxel_size=0.1):
        self._L = _lib.load()
        self._h = C.c_void_p()
        self.device = device
        rc = self._L.osep_map_create(device, capacity_points, max_frame_points, voxel_size, C.byref(self._h))
        if rc != 0:
            raise OsepError("osep_map_create failed (%d): %s" % (rc, _lib.last_error()))

    def close(self):
        if self._h:
            self._L.osep_map_destroy(self._h); self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _tf(self, translation, rotation_xyzw):
        t, q = _dbl(translation, 3), _dbl(rotation_xyzw, 4)
        return (t, q, t.ctypes.data_as(C.c_void_p) if t is not None else None, q.ctypes.data_as(C.c_void_p) if q is not None else None)

    def insert(self, xyz, translation=None, rotation_xyzw=None):
        """host frame (n, 3|4) float32; returns the number of new voxels"""
        a = np.ascontiguousarray(xyz, dtype=np.float32)
        t, q, tp, qp = self._tf(translation, rotation_xyzw)
        nn = C.c_int(0)
        rc = self._L.osep_map_insert(self._h, a.ctypes.data_as(C.c_void_p), a.shape[0], a.shape[1], tp, qp, C.byref(nn))
        if rc != 0:
            raise OsepError("osep_map_insert failed (%d): %s" % (rc, _lib.last_error()))
        return nn.value

    def insert_device(self, dev_ptr, n, stride_floats, translation=None, rotation_xyzw=None):
        t, q, tp, qp = self._tf(translation, rotation_xyzw)
        nn = C.c_int(0)
        rc = self._L.osep_map_insert_device(self._h, C.c_void_p(dev_ptr), n, stride_floats, tp, qp, C.byref(nn))
        if rc != 0:
            raise OsepError("osep_map_insert_device failed (%d): %s" % (rc, _lib.last_error()))
        return nn.value

    def points_device(self):
        """(device pointer, n): n records {x, y, z, 1}"""
        p = C.c_void_p(); n = C.c_int(0)
        self._L.osep_map_points_device(self._h, C.byref(p), C.byref(n))
        return p.value, n.value

    def points_tensor(self):
        """the dense point array as a CUDA tensor view (n, 4), no copy (for tiled.TileRunner)"""
        import torch
        ptr, n = self.points_device()

        class _Arr:
            __cuda_array_interface__ = {"shape": (max(n, 1), 4), "typestr": "<f4", "data": (ptr, False), "version": 2}
        t = torch.as_tensor(_Arr(), device=torch.device("cuda", self.device))
        return t[:n]

    def clear(self):
        self._L.osep_map_clear(self._h)
