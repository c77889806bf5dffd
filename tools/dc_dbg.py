"""Developer script: phase times of k_dcrosa_cluster (globaltimer stamps of CTA 0 thread 0, OSEP_FUSED_DETAIL=1)."""
import sys, os
os.environ.setdefault("OSEP_FUSED_DETAIL", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, ctypes as C
import osep_path_planner_b200 as osep
from osep_path_planner_b200 import fixtures as fx
P = fx.turbine()
r = osep.Rosa(capacity_points=P.shape[0])
r.set_input(P); r.set_graph(True, 3.0)
for k in range(3):
    r.rosa_run()
    buf = np.zeros(128, np.int64)
    r._L.osep_rosa_get(r._h, b"fused_dbg", buf.ctypes.data_as(C.c_void_p), buf.nbytes)
    t = buf[64:64 + 34]
    print("frame %d: stage %.1f us |" % (k, (t[1] - t[0]) / 1e3), " ".join("%.1f" % ((t[i + 1] - t[i]) / 1e3) for i in range(1, 22)), "| dcrosa %.1f" % ((t[22] - t[0]) / 1e3), "| vs:", " ".join("%.1f" % ((t[i + 1] - t[i]) / 1e3) for i in range(22, 31)), "| total %.1f" % ((t[31] - t[0]) / 1e3))
