"""A/B helper: mean device time of the fused drosa stage and of the whole frame over a few turbine100k frames
(profile mode for the stage, normal mode for the frame).  Env knobs of the library (OSEP_*) select the variant."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import osep_path_planner_b200 as osep
from osep_path_planner_b200 import fixtures as fx
P = fx.turbine()
r = osep.Rosa(capacity_points=P.shape[0])
r.set_input(P)
for _ in range(20): r.rosa_run()
fr = []
for _ in range(30): r.rosa_run(); fr.append(r.result.gpu_ms)
r.set_profile(True)
st = []
for _ in range(10): r.rosa_run(); st.append(r.stage_ms()["drosa_orient_smooth"])
print("%s frame %.4f ms (min %.4f)  drosa %.4f ms" % (os.environ.get("TAG", ""), float(np.mean(fr)), float(np.min(fr)), float(np.mean(st))))
