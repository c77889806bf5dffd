"""BASELINE.json configs[2]: streaming sequence of incremental 50k-point LiDAR-style frames (exploration
replay): Rosa per frame on the GPU -> analytic sensor->odom transform -> GSkel::gskel_run; per-frame latency
p50/p99 (host wall clock around the two reference-facing calls, H2D/D2H included).
Usage: python tools/stream_bench.py [n_frames] [points]   -> one JSON line"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import osep_path_planner_b200 as osep
from osep_path_planner_b200 import fixtures as fx, GSkel, GSkelConfig

n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
points = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
r = osep.Rosa(capacity_points=points)
g = GSkel(GSkelConfig(gnd_th=-100.0))      # fixtures sit below the default 60 m ground threshold
lat_rosa, lat_gskel, nv, published = [], [], [], 0
for f in range(n_frames):
    P, R, t = fx.stream_frame(f, n=points, n_frames=n_frames)
    buf = r.set_input(P)
    t0 = time.perf_counter()
    ok = r.rosa_run()
    t1 = time.perf_counter()
    if ok:
        v = np.ascontiguousarray(r.output_vertices())
        v_odom = (v.astype(np.float64) @ R.T + t).astype(np.float32)
        g.set_input(v_odom)
        t2 = time.perf_counter()
        published += int(g.gskel_run())
        t3 = time.perf_counter()
        lat_gskel.append(1e3 * (t3 - t2)); nv.append(len(v))
    if f >= 5:
        lat_rosa.append(1e3 * (t1 - t0))
lr, lg = np.array(lat_rosa), np.array(lat_gskel[5:])
print(json.dumps({"config": "stream of %d x %d-point frames (helix replay), 1 x B200" % (n_frames, points),
                  "rosa_ms": {"p50": float(np.percentile(lr, 50)), "p99": float(np.percentile(lr, 99)), "mean": float(lr.mean()), "max": float(lr.max())},
                  "gskel_ms": {"p50": float(np.percentile(lg, 50)), "p99": float(np.percentile(lg, 99)), "mean": float(lg.mean())},
                  "vertices_per_frame_mean": float(np.mean(nv)), "global_vertices": int(len(g.output_gskel())), "ticks_published": published}))
