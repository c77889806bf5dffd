"""Developer script: run K frames of one fixture through the host-buffer API (for ncu launch lists)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import osep_path_planner_b200 as osep
from osep_path_planner_b200 import fixtures as fx

name = sys.argv[1] if len(sys.argv) > 1 else "turbine100k"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 3
P = fx.turbine() if name == "turbine100k" else fx.cylinder()
r = osep.Rosa(capacity_points=P.shape[0])
r.set_input(P)
r.set_graph(True, 3.0)          # frame -> skeleton graph inside the frame, as bench.py runs it
for k in range(K):
    ok = r.rosa_run()
    print(k, ok, r.result.n_vertices, "gpu_ms %.3f" % r.result.gpu_ms)
