import sys, os
os.environ.setdefault("OSEP_FUSED_DETAIL", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, ctypes as C
import osep_path_planner_b200 as osep
from osep_path_planner_b200 import fixtures as fx
P = fx.turbine()
r = osep.Rosa(capacity_points=P.shape[0]); r.set_profile(True)
r.set_input(P)
for k in range(3):
    r.rosa_run()
buf = np.zeros(128, np.int64)
n = r._L.osep_rosa_get(r._h, b"fused_dbg", buf.ctypes.data_as(C.c_void_p), buf.nbytes)
d = buf[:64].reshape(8, 8)
t0 = d[:6, 0].min()
for it in range(6):
    b, f, e, steps, spins, M = d[it, :6]
    print("      QR iterations: mean per point %.2f, mean of per-step max %.2f" % ((int(M) >> 32) / 851.0, (int(M) & 0xffffffff) / max(1, steps)))
    print("      compute cycles/step %.0f  release cycles/step %.0f (of which before the solve %.0f)" % (d[it, 6] / max(1, steps), d[it, 7] / max(1, steps), -spins / max(1, steps)))
    print("it %d: begin +%.1f us first_step +%.1f us end +%.1f us | steps %d spins %d | busy per step %.2f us" % (it, (b - t0) / 1e3, (f - t0) / 1e3, (e - t0) / 1e3, steps, spins, (e - f) / 1e3 / max(1, steps)))
k0, k1 = int(d[7, 0]), int(d[7, 1])
print("kernel: entry %+.1f us, last CTA exit %+.1f us (relative to the first chain's walker start)" % ((k0 - t0) / 1e3, (k1 - t0) / 1e3))
print(r.stage_ms())

print("knn_stats [units, slow, overflow, retry, grow]", r.tap("knn_stats"))

