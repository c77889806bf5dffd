"""Per-point timeline of k_drosa_fused (OSEP_FUSED_DETAIL=1): where the lag between the Gauss-Seidel chains goes.
Stamps (globaltimer ns) per (iteration, point): 0 = G solved by the walker, 1 = task (it, p) published, 2 = worker start,
3 = O flag up, 4 = O staged in the chain CTA, 5 = active set done, 6 = eigen solve done."""
import sys, os
os.environ.setdefault("OSEP_FUSED_DETAIL", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, ctypes as C
import osep_path_planner_b200 as osep
from osep_path_planner_b200 import fixtures as fx

P = fx.turbine() if len(sys.argv) < 2 else fx.stream_frame(int(sys.argv[1]), n=50000, n_frames=1000)[0]
r = osep.Rosa(capacity_points=P.shape[0]); r.set_profile(True)
r.set_input(P)
for k in range(3):
    r.rosa_run()
Mcap = 3072
buf = np.zeros(Mcap * 9 * 16, np.int64)
n = r._L.osep_rosa_get(r._h, b"fused_tl", buf.ctypes.data_as(C.c_void_p), buf.nbytes)
tl = buf.reshape(9, Mcap, 16).astype(np.float64)
M = int(r.tap("active_size").shape[0])
surf = r.tap("surf_idx").reshape(-1, 10)[:M]
tl = tl[:, :M]
t0 = tl[0, :, 2][tl[0, :, 2] > 0].min()
us = lambda a: a / 1e3


def q(a):
    a = np.asarray(a)
    return "mean %7.2f  p50 %7.2f  p90 %7.2f  max %7.2f" % (a.mean(), np.percentile(a, 50), np.percentile(a, 90), a.max())


for it in (0, 3):
    T = tl[it]
    line = "it %d:" % it
    if it >= 1:
        print(line, "publish - G(it-1) solved  ", q(us(T[:, 1] - tl[it - 1][:, 0])))
        print(line, "worker start - publish    ", q(us(T[:, 2] - T[:, 1])))
    else:
        print(line, "worker start rel. t0      ", q(us(T[:, 2] - t0)))
    if it < 6:
        print(line, "operand loads + slab      ", q(us(T[:, 7] - T[:, 2])))
        print(line, "BFS                       ", q(us(T[:, 5] - T[:, 7])))
        print(line, "sums + eigen              ", q(us(T[:, 6] - T[:, 5])))
        print(line, "sums + eigen: SM cycles per ns", q((T[:, 9] - T[:, 8]) / np.maximum(T[:, 6] - T[:, 5], 1.0)), " cycles", q(T[:, 9] - T[:, 8]))
        print(line, "variance + flag           ", q(us(T[:, 3] - T[:, 6])))
        print(line, "task total                ", q(us(T[:, 3] - T[:, 2])))
        print(line, "staged - flag             ", q(us(T[:, 4] - T[:, 3])))
        if it >= 1:
            print(line, "O(it) staged - G(it-1) solved", q(us(T[:, 4] - tl[it - 1][:, 0])))
        # slack of the walker: solved - last dependency satisfied
        dep = T[:, 4].copy()
        for p in range(M):
            js = surf[p]
            d = max(T[p, 4], T[js, 4].max())
            lo = js[js < p]
            if len(lo):
                d = max(d, T[lo, 0].max())
            dep[p] = d
        print(line, "G solved - last dependency ", q(us(T[:, 0] - dep)))
        print(line, "   first G %.1f us, last G %.1f us (rel. first worker start)" % (us(T[:, 0].min() - t0), us(T[:, 0].max() - t0)))
print(r.stage_ms())
