"""Developer script: kNN fall-back counters of a frame (knn_stats tap) for a few clouds."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, ctypes as C
import osep_path_planner_b200 as osep
from osep_path_planner_b200 import fixtures as fx
for name, P in (("turbine100k", fx.turbine()), ("cyl20k", fx.cylinder()), ("stream0", fx.stream_frame(0)[0]), ("stream7", fx.stream_frame(7)[0])):
    r = osep.Rosa(capacity_points=P.shape[0])
    r.set_input(P)
    for k in range(2):
        r.rosa_run()
        buf = np.zeros(8, np.int32)
        r._L.osep_rosa_get(r._h, b"knn_stats", buf.ctypes.data_as(C.c_void_p), buf.nbytes)
        print(name, k, "n", P.shape[0], "units %d slow %d over %d retry %d grow %d" % tuple(buf[:5]), "gpu_ms %.3f" % r.result.gpu_ms)
