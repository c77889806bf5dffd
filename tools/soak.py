"""Developer script: race soak.  The same frame is run many times through the production path (CUDA graph replay,
two streams, persistent cooperative drosa kernel); every run must reproduce the first run's vertices and frame
summary bit for bit.  usage: python tools/soak.py [iterations]"""
import sys, os, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import osep_path_planner_b200 as osep
from osep_path_planner_b200 import fixtures as fx

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
bad = 0
for name, P in (("turbine100k", fx.turbine()), ("cyl20k", fx.cylinder()), ("stream50k_f300", fx.stream_frame(300, n=50000)[0]), ("turbine30k", fx.turbine(30000, seed=5))):
    r = osep.Rosa(capacity_points=len(P))
    r.set_input(P)
    r.set_graph(True, 3.0)                 # the frame's tail builds the skeleton graph as well (graph_adj + Boruvka mst)
    ref = None; mism = 0
    for k in range(iters):
        assert r.rosa_run()
        g = r.graph(3.0)
        h = hashlib.sha1(np.ascontiguousarray(r.output_vertices()).tobytes() + np.ascontiguousarray(g["edges"]).tobytes() + np.ascontiguousarray(g["edge_w"]).tobytes()
                         + np.ascontiguousarray(g["rng_idx"]).tobytes()).hexdigest() + "|%d|%d|%d" % (r.result.n_downsampled, r.result.n_simi, r.result.flags)
        if ref is None: ref = h
        elif h != ref: mism += 1
    print("%-16s %d runs, %d mismatches, V=%d M=%d" % (name, iters, mism, r.result.n_vertices, r.result.n_downsampled))
    bad += mism
    r.close()
# batched frames in flight on several streams (BASELINE configs[3]): every frame must equal its single-frame result
frames = [fx.batch_frame(b, n=40000) for b in range(12)]
r = osep.Rosa(capacity_points=40000)
single = []
for P in frames:
    r.set_input(P); assert r.rosa_run(); single.append(np.ascontiguousarray(r.output_vertices()).copy())
r.set_lanes(5)
mism = 0; rounds = max(1, iters // 100)
for k in range(rounds):
    order = [(b + k) % len(frames) for b in range(len(frames) * 2)]
    res, outs = r.run_batch([fx.to_xyz4(frames[b]) for b in order])
    for b, v in zip(order, outs):
        if not np.array_equal(np.ascontiguousarray(v).view(np.uint32), single[b].view(np.uint32)): mism += 1
print("batch            %d rounds x %d frames on 5 lanes, %d mismatches" % (rounds, 2 * len(frames), mism))
bad += mism
r.close()
print("SOAK", "OK" if bad == 0 else "FAILED")
sys.exit(1 if bad else 0)
