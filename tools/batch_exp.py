"""Developer script: batched throughput (64 distinct device-resident 100k frames) for lane counts / ablated configs.
usage: batch_exp.py [lanes ...]   env ABL=drosa|post|none"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import osep_path_planner_b200 as osep
from osep_path_planner_b200 import fixtures as fx
abl = os.environ.get("ABL", "none")
cfg = osep.RosaConfig()
if abl in ("drosa", "post"): cfg.niter_drosa = 0
if abl == "post": cfg.niter_dcrosa = 0
B = 96
frames = [torch.from_numpy(fx.to_xyz4(fx.batch_frame(b))).cuda() for b in range(16)]
ptrs = [frames[b % 16].data_ptr() for b in range(B)]; cnts = [100000] * B
for lanes in [int(a) for a in sys.argv[1:]] or [12]:
    r = osep.Rosa(cfg=cfg, capacity_points=100000)
    if os.environ.get('PRE'):
        r.set_graph(True, 3.0)
        for _ in range(20): r.run_device(ptrs[0], 100000, 4); r.finish()
    r.set_lanes(lanes)
    r.run_batch(ptrs, device_ptrs=True, counts=cnts)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); r.run_batch(ptrs, device_ptrs=True, counts=cnts); ts.append(time.perf_counter() - t0)
    print("abl=%s lanes %2d: %.0f frames/s (runs %s)" % (abl, lanes, B / np.median(ts), " ".join("%.0f" % (B / t) for t in ts)), flush=True)
    r.close()
