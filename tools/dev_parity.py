"""Developer script: run one frame through the CUDA path and the CPU oracle and compare every
tap bit-for-bit.  Usage: python tools/dev_parity.py [cyl20k|turbine100k|...]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import pyoracle as po
import osep_path_planner_b200 as osep
from osep_path_planner_b200 import fixtures as fx


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32) if a.dtype == np.float32 else a


def cmp(name, a, b):
    a = np.asarray(a); b = np.asarray(b)
    if a.shape != b.shape:
        print("  %-16s SHAPE MISMATCH gpu %s oracle %s" % (name, a.shape, b.shape)); return False
    if a.size == 0:
        print("  %-16s ok (empty)" % name); return True
    eq = bits(a) == bits(b)
    if a.dtype == np.float32:
        eq = eq | (np.isnan(a) & np.isnan(b))
    nbad = int((~eq).sum())
    if nbad == 0:
        print("  %-16s ok  %s" % (name, a.shape)); return True
    idx = np.argwhere(~eq)[0]
    extra = ""
    if a.dtype == np.float32:
        with np.errstate(invalid="ignore"):
            extra = " maxabs %.3e" % float(np.nanmax(np.abs(a.astype(np.float64) - b.astype(np.float64))))
    print("  %-16s MISMATCH %d/%d first at %s gpu %s oracle %s%s" % (name, nbad, a.size, tuple(idx), a[tuple(idx)], b[tuple(idx)], extra))
    return False


def run(name, P):
    print("==", name, P.shape)
    o = po.RosaOracle()
    t0 = time.time(); ok_o = o.run(P); t_cpu = time.time() - t0
    r = osep.Rosa(capacity_points=max(1024, P.shape[0]))
    r.set_debug(True)
    r.set_input(P)
    ok_g = r.rosa_run()
    res = r.result
    print("  oracle ok=%s stage=%d  cpu %.1f ms | gpu ok=%s status=%d flags=%d gpu_ms=%.3f" % (ok_o, o.failed_stage, t_cpu * 1e3, ok_g, res.status, res.flags, res.gpu_ms))
    print("  N'=%d M=%d V=%d passes=%d S=%d leaf=%.9g used=%.9g | oracle leaf=%.9g used=%.9g flags=%d" % (
        res.n_filtered, res.n_downsampled, res.n_vertices, res.n_passes, res.n_simi, res.leaf_size, res.leaf_used, o.leaf_size, o.leaf_used, o.flags))
    allok = True
    for tap in ["filt_idx", "filtered", "leaf_hist", "nvox_hist", "knn_full", "normals_full", "vox_keys", "vox_count", "pts", "nrms",
                "surf_idx", "simi_off", "simi_idx", "vset_iters", "active_size", "poor_idx", "pset_drosa", "pset", "seeds", "corresp",
                "skelver_sampled", "skelver"]:
        try:
            a = r.tap(tap); b = o.tap(tap)
        except KeyError as e:
            print("  %-16s missing %s" % (tap, e)); continue
        if tap == "knn_full":
            a = a.reshape(-1); b = b.reshape(-1)
        allok &= cmp(tap, a, b)
    v = r.output_vertices()
    allok &= cmp("output_vertices", np.ascontiguousarray(v), o.tap("skelver"))
    # timing: a few repeated runs
    ts = []
    for _ in range(5):
        r.rosa_run(); ts.append(r.result.gpu_ms)
    print("  gpu_ms runs:", ["%.3f" % t for t in ts])
    g = r.graph(3.0)
    go = po.GraphOracle(o.tap("skelver"), 3.0)
    allok &= cmp("mst_edges", g["edges"].reshape(-1), go.tap("mst_edges"))
    allok &= cmp("adj_idx", g["adj_idx"], go.tap("adj_idx"))
    allok &= cmp("rng_idx", g["rng_idx"], go.tap("rng_idx"))
    allok &= cmp("type", g["type"], go.tap("type"))
    print("  ALL OK" if allok else "  SOME MISMATCH")
    return allok


if __name__ == "__main__":
    which = sys.argv[1:] or ["cyl20k", "turbine100k"]
    ok = True
    for w in which:
        if w == "cyl20k": ok &= run(w, fx.cylinder())
        elif w == "turbine100k": ok &= run(w, fx.turbine())
        elif w == "cyl5k": ok &= run(w, fx.cylinder(5000, seed=3))
        elif w.startswith("batch"): ok &= run(w, fx.batch_frame(int(w[5:])))
    sys.exit(0 if ok else 1)
