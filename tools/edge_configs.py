"""Developer script: parity against the oracle at the edges of the configuration space (largest M, tiny M, iteration counts, k)."""
import sys; sys.path.insert(0, "/root/repo")
import numpy as np, time
import osep_path_planner_b200 as osep
from osep_path_planner_b200 import fixtures as fx
from oracle import pyoracle as po
for kw, P in ((dict(max_points=1500), fx.turbine(60000, seed=3)), (dict(max_points=1536), fx.cylinder(40000, seed=4)), (dict(max_points=120, min_points=50), fx.cylinder(3000, seed=5)),
              (dict(niter_drosa=8), fx.cylinder(20000, seed=7)), (dict(niter_drosa=1, niter_dcrosa=1, niter_smooth=0), fx.cylinder(20000, seed=8)),
              (dict(nb_knn=20, ne_knn=40), fx.turbine(30000, seed=9))):
    o = po.RosaOracle(po.RosaConfig.default(**kw)); ok_o = o.run(P)
    r = osep.Rosa(osep.RosaConfig(**kw), capacity_points=len(P))
    r.set_input(P)
    t0 = time.perf_counter(); ok_g = r.rosa_run(); ok_g = r.rosa_run(); ok_g = r.rosa_run(); dt = (time.perf_counter() - t0) / 3
    same = ok_o == ok_g and (not ok_o or np.array_equal(np.ascontiguousarray(r.output_vertices()).view(np.uint32), o.tap("skelver").view(np.uint32)))
    print(kw, "oracle ok", ok_o, "stage", o.failed_stage, "| gpu ok", ok_g, "status", r.result.status, "M", r.result.n_downsampled, "V", r.result.n_vertices, "gpu_ms %.2f" % r.result.gpu_ms, "| bit-exact:", same)
    r.close()
