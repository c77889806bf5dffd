"""Developer script: phase times of k_frame_tail (globaltimer stamps, OSEP_FUSED_DETAIL=1)."""
import sys, os
os.environ.setdefault("OSEP_FUSED_DETAIL", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, ctypes as C
import osep_path_planner_b200 as osep
from osep_path_planner_b200 import fixtures as fx
P = fx.turbine()
r = osep.Rosa(capacity_points=P.shape[0])
r.set_input(P); r.set_graph(True, 3.0)
for k in range(4):
    r.rosa_run()
    buf = np.zeros(128, np.int64)
    r._L.osep_rosa_get(r._h, b"fused_dbg", buf.ctypes.data_as(C.c_void_p), buf.nbytes)
    t = buf[48:56]
    print("frame %d V=%d: smooth %.1f us | graph_adj %.1f | mst keys %.1f sort %.1f kruskal %.1f | export %.1f | total %.1f" % (
        k, r.result.n_vertices, (t[1] - t[0]) / 1e3, (t[2] - t[1]) / 1e3, (t[4] - t[2]) / 1e3, (t[5] - t[4]) / 1e3, (t[3] - t[5]) / 1e3, (t[7] - t[3]) / 1e3, (t[7] - t[0]) / 1e3))
    print("   smooth: compaction %.1f lists %.1f iterations %s (single-pass lists; the two-pass path has one stamp more)" % ((buf[56] - t[0]) / 1e3, (buf[57] - buf[56]) / 1e3, " ".join("%.1f" % ((buf[58 + i] - buf[57 + i]) / 1e3) for i in range(3))))
