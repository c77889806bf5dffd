"""Developer script: one large frame (default 5M points, 4-turbine scene) -> stage times + kNN statistics."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import osep_path_planner_b200 as osep
from osep_path_planner_b200 import fixtures as fx
total = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
parts = [fx.turbine(total // 4, seed=100 + k) + np.array([ox, oy, 0], np.float32) for k, (ox, oy) in enumerate([(0, 0), (60, 0), (0, 60), (60, 60)])]
P = np.concatenate(parts).astype(np.float32)
r = osep.Rosa(osep.RosaConfig(pts_dist_lim=1.0e4), capacity_points=len(P) + 1024)
r.set_profile(True)
r.set_input(P)
for k in range(3):
    ok = r.rosa_run()
    print(k, ok, r.result.status, "M", r.result.n_downsampled, "gpu_ms %.2f" % r.result.gpu_ms)
print({k: round(v, 3) for k, v in r.stage_ms().items()})
print("knn_stats [units, slow, overflow, retry, grow]", r.tap("knn_stats"))
