"""Developer script: what bounds batched frames?  Device-resident 100k frames through osep_rosa_run_batch_device with the
work cut away stage by stage:  FLOOR = frames that fail the min_points gate right after the filter (every later kernel of
the frame graph exits at once: what is left is launch / graph / lane overhead);  PRE = niter_drosa = niter_dcrosa = 0
(N-scale stages + similarity + cover only);  FULL.   usage: batch_floor.py [lanes ...]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import osep_path_planner_b200 as osep
from osep_path_planner_b200 import fixtures as fx
B = 96
frames = [torch.from_numpy(fx.to_xyz4(fx.batch_frame(b))).cuda() for b in range(16)]
ptrs = [frames[b % 16].data_ptr() for b in range(B)]; cnts = [100000] * B
for mode in ["FLOOR", "PRE", "FULL"]:
    cfg = osep.RosaConfig()
    if mode == "FLOOR": cfg.min_points = 10 ** 9
    if mode == "PRE": cfg.niter_drosa = 0; cfg.niter_dcrosa = 0
    for lanes in [int(a) for a in sys.argv[1:]] or [16]:
        r = osep.Rosa(cfg=cfg, capacity_points=100000)
        r.set_lanes(lanes)
        r.run_batch(ptrs, device_ptrs=True, counts=cnts)
        ts = []
        for _ in range(3):
            t0 = time.perf_counter(); res, _ = r.run_batch(ptrs, device_ptrs=True, counts=cnts); ts.append(time.perf_counter() - t0)
        print("%-5s lanes %2d: %6.0f frames/s  %.3f ms/frame (status of frame 0: %d)" % (mode, lanes, B / np.median(ts), 1e3 * np.median(ts) / B, res[0].status), flush=True)
        r.close()
