"""Developer script: the bench's varying-n stream (bench.py `stream`), slowest frames with their device time, M, passes."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ctypes as C
import osep_path_planner_b200 as osep
from osep_path_planner_b200 import fixtures as fx
from osep_path_planner_b200.rosa import RosaResult
n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rng = np.random.default_rng(7)
sf = []
for f in range(n_frames):
    P, Rm, tv = fx.stream_frame(f * 5, n=50000)
    keep_n = int(rng.integers(42000, 50001))
    sf.append(torch.from_numpy(fx.to_xyz4(P[:keep_n])).pin_memory())
r = osep.Rosa(capacity_points=50000); r.set_graph(True, 3.0)
L = r._L; res = RosaResult()
for wu in range(3):
    L.osep_rosa_run(r._h, C.c_void_p(sf[wu].data_ptr()), int(sf[wu].shape[0]), 4, C.byref(res))
rows = []
for i, buf in enumerate(sf):
    t0 = time.perf_counter()
    rc = L.osep_rosa_run(r._h, C.c_void_p(buf.data_ptr()), int(buf.shape[0]), 4, C.byref(res))
    t1 = time.perf_counter()
    rows.append((1e3 * (t1 - t0), res.gpu_ms, i, int(buf.shape[0]), res.n_downsampled, res.n_vertices, res.n_passes, res.n_simi, rc))
rows.sort(key=lambda x: -x[0])
print("slowest: (wall_ms, gpu_ms, frame, n, M, V, passes, S, rc)")
for x in rows[:12]: print("  %.3f %.3f %s" % (x[0], x[1], x[2:]))
w = np.array([x[0] for x in rows]); g = np.array([x[1] for x in rows])
print("wall p50 %.3f p90 %.3f p99 %.3f max %.3f | device p50 %.3f p90 %.3f p99 %.3f max %.3f" % (*np.percentile(w, [50, 90, 99]), w.max(), *np.percentile(g, [50, 90, 99]), g.max()))
