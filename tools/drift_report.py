"""Report-level comparison of the two oracle modes (SURVEY.md 8d "parity gates", last row): canonical (what the CUDA
path is bit-compared with) versus faithful (the reference's own traversal/summation orders, PCL's trigonometric
eigen33, libm cbrt).  The pipeline is chaotic at ulp scale, so these numbers are reported, not gated:
vertex-set Hausdorff and chamfer distance in units of leaf, |dV|, Jaccard index of the skeleton-graph edge sets
(edges matched through nearest vertices), per fixture.  CPU only.   usage: python tools/drift_report.py > profiles/..."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import pyoracle as po
from osep_path_planner_b200 import fixtures as fx


def edge_set(g, remap=None):
    e = g.tap("mst_edges").reshape(-1, 2)
    if remap is not None:
        e = remap[e]
    return {(min(a, b), max(a, b)) for a, b in e.tolist() if a != b}


rows = []
for name, P in (("cyl20k", fx.cylinder()), ("cyl50k", fx.cylinder(50000, seed=1)), ("turbine100k", fx.turbine()), ("turbine100k_seed1", fx.turbine(seed=1))):
    a = po.RosaOracle(); b = po.RosaOracle(faithful_normals=1, faithful_order=1, libm_cbrt=1)
    assert a.run(P) and b.run(P)
    va, vb = a.tap("skelver"), b.tap("skelver")
    d = np.sqrt(((va[:, None, :].astype(np.float64) - vb[None, :, :]) ** 2).sum(-1))
    leaf = float(a.leaf_size)
    ga, gb = po.GraphOracle(va, 3.0), po.GraphOracle(vb, 3.0)
    ea = edge_set(ga); eb = edge_set(gb, remap=d.argmin(0))          # faithful vertices mapped to their nearest canonical vertex
    na, nb = a.tap("normals_full"), b.tap("normals_full")
    ang = np.degrees(np.arccos(np.clip(np.abs((na * nb).sum(1)), 0, 1)))
    rows.append({"fixture": name, "M_canonical": int(len(a.tap("pts"))), "M_faithful": int(len(b.tap("pts"))),
                 "voxel_keys_equal": bool(np.array_equal(a.tap("vox_keys"), b.tap("vox_keys"))),
                 "knn_sets_equal": bool(np.array_equal(np.sort(a.tap("knn_full").reshape(len(na), -1), 1), np.sort(b.tap("knn_full").reshape(len(nb), -1), 1))),
                 "normal_angle_deg": {"median": float(np.nanmedian(ang)), "p99": float(np.nanpercentile(ang, 99)), "max": float(np.nanmax(ang))},
                 "simi_lists_equal": bool(np.array_equal(a.tap("simi_idx"), b.tap("simi_idx"))),
                 "V_canonical": int(len(va)), "V_faithful": int(len(vb)), "leaf": leaf,
                 "hausdorff_over_leaf": float(max(d.min(1).max(), d.min(0).max()) / leaf),
                 "chamfer_over_leaf": float(0.5 * (d.min(1).mean() + d.min(0).mean()) / leaf),
                 "edge_jaccard": len(ea & eb) / max(1, len(ea | eb))})
print(json.dumps({"what": "canonical vs faithful oracle mode (reported, not gated; SURVEY.md 8d / H1)", "rows": rows}, indent=1))
