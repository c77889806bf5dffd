"""Developer script: per-frame device time, M, V and kNN fallback counters over a slice of the C3 stream."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import osep_path_planner_b200 as osep
from osep_path_planner_b200 import fixtures as fx
n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 100
step = int(sys.argv[2]) if len(sys.argv) > 2 else 10
r = osep.Rosa(capacity_points=50000)
rows = []
for f in range(0, n_frames * step, step):
    P, R, t = fx.stream_frame(f, n=50000, n_frames=1000)
    r.set_input(P)
    r.rosa_run(); r.rosa_run()
    ok = r.rosa_run()
    st = r.tap("knn_stats")
    rows.append((f, r.result.gpu_ms, r.result.n_downsampled, r.result.n_vertices, r.result.n_passes, int(st[1]), int(st[2])))
rows.sort(key=lambda x: -x[1])
print("slowest frames: (frame, gpu_ms, M, V, passes, knn_slow, knn_overflow)")
for x in rows[:12]: print(x)
a = np.array([x[1] for x in rows])
print("gpu_ms p50 %.3f p90 %.3f p99 %.3f max %.3f; frames with overflow %d, with slow %d" % (np.percentile(a, 50), np.percentile(a, 90), np.percentile(a, 99), a.max(), sum(1 for x in rows if x[6] > 0), sum(1 for x in rows if x[5] > 0)))
