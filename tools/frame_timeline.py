"""Developer script: where a frame's time goes -- globaltimer stamps at the entry of selected kernels (tap "timeline"),
taken in a normal run (streams overlapped, CUDA graph replay).  usage: frame_timeline.py [turbine100k|cyl20k] [frames]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, ctypes as C
import osep_path_planner_b200 as osep
from osep_path_planner_b200 import fixtures as fx
NAMES = ["filter", "leaf pass 0", "kNN grid", "kNN scan", "normals", "voxel final", "reduce<0>", "surf kNN", "schedule", "schedule end", "reduce<1> (join)", "similarity",
         "drosa", "fix-up", "dcrosa+vs", "tail", "end"]
name = sys.argv[1] if len(sys.argv) > 1 else "turbine100k"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 6
P = fx.turbine() if name == "turbine100k" else fx.cylinder()
r = osep.Rosa(capacity_points=P.shape[0])
r.set_input(P); r.set_graph(True, 3.0)
rows = []
for k in range(K):
    r.rosa_run()
    buf = np.zeros(18, np.uint64)
    r._L.osep_rosa_get(r._h, b"timeline", buf.ctypes.data_as(C.c_void_p), buf.nbytes)
    t = buf.astype(np.int64)
    rows.append((t[:17] - t[0]) / 1e3)
med = np.median(np.array(rows[2:]), axis=0)
print("%s: kernel entry times in us after the filter's scan kernel (median of frames 2..%d), gpu_ms %.3f" % (name, K - 1, r.result.gpu_ms))
for n, v in sorted(zip(NAMES, med), key=lambda x: x[1]):
    print("  %8.1f  %s" % (v, n))
