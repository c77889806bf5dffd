"""Developer script: stage times + fused-kernel chain records of one frame of the bench's varying-n stream."""
import sys, os
os.environ.setdefault("OSEP_FUSED_DETAIL", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, ctypes as C
import osep_path_planner_b200 as osep
from osep_path_planner_b200 import fixtures as fx
want = [int(a) for a in sys.argv[1:]] or [75]
rng = np.random.default_rng(7)
keep = [int(rng.integers(42000, 50001)) for _ in range(max(want) + 1)]
for f in want:
    P = fx.stream_frame(f * 5, n=50000)[0][:keep[f]]
    r = osep.Rosa(capacity_points=50000); r.set_profile(True); r.set_input(P)
    for k in range(3): r.rosa_run()
    buf = np.zeros(128, np.int64)
    r._L.osep_rosa_get(r._h, b"fused_dbg", buf.ctypes.data_as(C.c_void_p), buf.nbytes)
    d = buf[:64].reshape(8, 8); t0 = d[:6, 0].min()
    res = r.result
    print("frame %d n %d M %d V %d passes %d S %d flags 0x%x gpu_ms %.3f" % (f, P.shape[0], res.n_downsampled, res.n_vertices, res.n_passes, res.n_simi, res.flags, res.gpu_ms))
    print("  ", {k: round(v * 1e3, 1) for k, v in r.stage_ms().items()})
    for it in range(6):
        b, fs, e, steps, spins = d[it, :5]
        print("   chain %d: first step +%.1f end +%.1f us, steps %d, %.2f us per step" % (it, (fs - t0) / 1e3, (e - t0) / 1e3, steps, (e - fs) / 1e3 / max(1, steps)))
    lv = r.tap("levels"); print("   DAG depth", int(lv.max()) if len(lv) else 0, "active_size mean %.1f max %d" % (r.tap("active_size").mean(), r.tap("active_size").max()))
    print("   knn_stats [units, slow(exact), overflow, retry, grow]", r.tap("knn_stats"))
    r.close()
