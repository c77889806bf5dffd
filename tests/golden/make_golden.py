"""Generates tests/golden/*.npz from the CPU oracle (canonical mode).

The reference (/root/reference) cannot be compiled or imported in this container (C++ against
PCL / Eigen / FLANN / ROS 2, none present) and ships no golden vectors of its own, so these
fixtures pin the ORACLE's outputs (regression anchors for both the oracle and the CUDA path).
The independent known-answer numbers the tests also check (N', L, M, leaf, S for cyl20k and
turbine100k) come from SURVEY.md section 8's sizing emulation, not from this script.
Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po                      # noqa: E402
from osep_path_planner_b200 import fixtures as fx      # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def dump(name, P, cfg=None):
    o = po.RosaOracle(cfg)
    ok = o.run(P)
    off, idx = o.csr("simi")
    out = dict(ok=ok, failed_stage=o.failed_stage, flags=o.flags, leaf_size=np.float32(o.leaf_size), leaf_used=np.float32(o.leaf_used),
               n_filtered=len(o.tap("filt_idx")), leaf_hist=o.tap("leaf_hist"), nvox_hist=o.tap("nvox_hist"),
               vox_keys=o.tap("vox_keys"), vox_count=o.tap("vox_count"), pts=o.tap("pts"), nrms=o.tap("nrms"),
               simi_off=off, simi_idx_sum=np.int64(idx.astype(np.int64).sum()), simi_n=len(idx),
               surf_idx=o.tap("surf_idx"), vset=o.tap("vset"), pset=o.tap("pset"), poor_idx=o.tap("poor_idx"),
               seeds=o.tap("seeds"), corresp=o.tap("corresp"), skelver=o.tap("skelver"))
    g = po.GraphOracle(o.tap("skelver"), 3.0)
    out["mst_edges"] = g.tap("mst_edges")
    out["graph_type"] = g.tap("type")
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "ok", ok, "M", len(out["pts"]), "V", len(out["skelver"]), "S", out["simi_n"])


if __name__ == "__main__":
    dump("cyl5k_seed3", fx.cylinder(5000, seed=3))
    dump("cyl20k_seed0", fx.cylinder())
    dump("cyl3k_k10_m300", fx.cylinder(3000, seed=7), po.RosaConfig.default(max_points=300, ne_knn=10, nb_knn=6))
