import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np
import osep_path_planner_b200 as osep
from osep_path_planner_b200 import fixtures as fx
P = fx.turbine()
r = osep.Rosa(capacity_points=P.shape[0]); r.set_profile(True)
r.set_input(P)
acc = None
for k in range(6):
    r.rosa_run()
    if k >= 2:
        s = r.tap("stage_ms"); acc = s if acc is None else acc + s
acc /= 4
print("ppc", os.environ.get("OSEP_KNN_PPC"), "grid %.4f knn %.4f" % (acc[2], acc[3]), "stats", r.tap("knn_stats"))
