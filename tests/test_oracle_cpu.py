"""CPU tests of the oracle (no GPU): primitives against numpy/libm, the pipeline against the
committed golden fixtures and against the independent sizing numbers of SURVEY.md section 8."""
import os

import numpy as np
import pytest

from oracle import pyoracle as po
from osep_path_planner_b200 import fixtures as fx
from conftest import assert_bit_equal

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_eig3_matches_numpy():
    rng = np.random.default_rng(0)
    for _ in range(300):
        B = rng.normal(size=(3, 3)).astype(np.float32)
        A = (B @ B.T).astype(np.float32)
        ok, ev, Q = po.eig3(A)
        assert ok
        w, _ = np.linalg.eigh(A.astype(np.float64))
        assert np.allclose(ev, w, rtol=2e-5, atol=2e-5 * max(1.0, abs(w).max()))
        assert ev[0] <= ev[1] <= ev[2]
        # columns are eigenvectors: A q = lambda q (float32 tolerance), orthonormal
        for c in range(3):
            assert np.allclose(A @ Q[:, c], ev[c] * Q[:, c], atol=5e-5 * max(1.0, abs(w).max()))
        assert np.allclose(Q.T @ Q, np.eye(3), atol=1e-5)


def test_eig3_degenerate_inputs():
    ok, ev, Q = po.eig3(np.zeros((3, 3), np.float32))
    assert ok and np.all(ev == 0) and np.array_equal(Q, np.eye(3, dtype=np.float32))
    ok, ev, Q = po.eig3(np.diag([3.0, 1.0, 2.0]).astype(np.float32))
    assert ok and np.array_equal(ev, np.array([1, 2, 3], np.float32))
    # only the lower triangle is read (Eigen semantics)
    A = np.array([[2, 9, 9], [1, 3, 9], [0.5, 0.25, 4]], np.float32)
    As = np.tril(A) + np.tril(A, -1).T
    assert_bit_equal(po.eig3(A)[1], po.eig3(As)[1], "lower-triangle read")


def test_pcl_eigen33_agrees_with_iterative_solver_within_tolerance():
    rng = np.random.default_rng(1)
    for _ in range(100):
        pts = rng.normal(size=(20, 3)) * np.array([1.0, 0.7, 0.05])
        C = np.cov(pts.T, bias=True).astype(np.float32)
        ev, v = po.pcl_eigen33(C)
        ok, evs, Q = po.eig3(C)
        assert abs(ev - evs[0]) < 1e-4 * max(1.0, evs[2])
        assert abs(abs(float(v @ Q[:, 0])) - 1.0) < 1e-3      # same direction up to sign


def test_canonical_cbrt_is_correctly_rounded_vs_libm():
    rng = np.random.default_rng(2)
    xs = np.concatenate([rng.uniform(1e-6, 1e6, 20000), 10.0 ** rng.uniform(-30, 30, 5000)]).astype(np.float32)
    mism = 0
    for x in xs:
        ref = np.float32(np.cbrt(np.float64(x)))          # correctly rounded (double cbrt then one rounding)
        if po.cbrtf(float(x)) != float(ref):
            mism += 1
    assert mism == 0, "%d of %d canonical cube roots are not correctly rounded" % (mism, len(xs))
    assert po.cbrtf(0.0) == 0.0 and po.cbrtf(27.0) == 3.0 and po.cbrtf(-8.0) == -2.0 and np.isnan(po.cbrtf(float("nan")))


def test_llt_and_ldlt_solve():
    rng = np.random.default_rng(3)
    for _ in range(100):
        B = rng.normal(size=(3, 3)); A = (B @ B.T + 0.1 * np.eye(3)).astype(np.float32)
        b = rng.normal(size=3).astype(np.float32)
        x_ref = np.linalg.solve(A.astype(np.float64), b.astype(np.float64))
        for kind in ("llt", "ldlt"):
            ok, x = po.solve3(A, b, kind)
            assert ok and np.allclose(x, x_ref, rtol=2e-3, atol=2e-3)
    ok, _ = po.solve3(np.diag([1.0, -1.0, 1.0]).astype(np.float32), np.ones(3, np.float32), "llt")
    assert not ok          # non-positive pivot -> NumericalIssue -> LDLT fallback in the reference (rosa.cpp:778-785)
    ok, x = po.solve3(np.diag([1.0, -2.0, 4.0]).astype(np.float32), np.ones(3, np.float32), "ldlt")
    assert ok and np.allclose(x, [1.0, -0.5, 0.25])


def test_kdtree_knn_is_exact_with_canonical_ties():
    rng = np.random.default_rng(4)
    P = rng.normal(size=(700, 3)).astype(np.float32)
    P[100:120] = P[50:70]                                  # exact duplicates -> distance ties
    P = np.round(P * 4) / 4                                # lattice -> many equal distances
    got = po.knn(P, 12)
    # brute force in float32 with the same operation order as FLANN L2_Simple
    for i in range(0, 700, 7):
        d = P[i][None, :] - P
        d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
        order = np.lexsort((np.arange(700), d2))[:12]
        assert np.array_equal(got[i], order)


@pytest.mark.parametrize("name,P,cfgkw", [
    ("cyl5k_seed3", lambda: fx.cylinder(5000, seed=3), {}),
    ("cyl20k_seed0", lambda: fx.cylinder(), {}),
    ("cyl3k_k10_m300", lambda: fx.cylinder(3000, seed=7), dict(max_points=300, ne_knn=10, nb_knn=6)),
])
def test_oracle_matches_golden(name, P, cfgkw):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    o = po.RosaOracle(po.RosaConfig.default(**cfgkw))
    ok = o.run(P())
    assert ok == bool(g["ok"]) and o.failed_stage == int(g["failed_stage"]) and o.flags == int(g["flags"])
    assert np.float32(o.leaf_size) == g["leaf_size"] and np.float32(o.leaf_used) == g["leaf_used"]
    for tap in ["leaf_hist", "nvox_hist", "vox_keys", "vox_count", "pts", "nrms", "surf_idx", "vset", "pset", "poor_idx",
                "seeds", "corresp", "skelver"]:
        assert_bit_equal(o.tap(tap), g[tap], tap)
    off, idx = o.csr("simi")
    assert_bit_equal(off, g["simi_off"], "simi_off")
    assert int(idx.astype(np.int64).sum()) == int(g["simi_idx_sum"])
    gr = po.GraphOracle(o.tap("skelver"), 3.0)
    assert_bit_equal(gr.tap("mst_edges"), g["mst_edges"], "mst_edges")


def test_known_answers_from_survey_sizing():
    """SURVEY.md section 8 sizing table (an independent numpy/scipy emulation of rosa.cpp):
    cyl20k: N'=20000, L=5, M=973, leaf 0.675, S=61.6k.  turbine100k: N'=98762, L=8 (no
    convergence), M=851, leaf 0.936 used / 0.887 reported, S=44.6k."""
    o = po.RosaOracle(stop_after_stage=3)
    assert o.run(fx.cylinder())
    assert len(o.tap("filt_idx")) == 20000 and len(o.tap("leaf_hist")) == 5 and len(o.tap("pts")) == 973
    assert abs(o.leaf_size - 0.675) < 5e-4 and abs(len(o.tap("simi_idx")) - 61600) < 100
    o = po.RosaOracle(stop_after_stage=3)
    assert o.run(fx.turbine())
    assert len(o.tap("filt_idx")) == 98762 and len(o.tap("leaf_hist")) == 8 and len(o.tap("pts")) == 851
    assert abs(o.leaf_used - 0.936) < 1e-3 and abs(o.leaf_size - 0.887) < 1e-3 and (o.flags & 1)
    assert abs(len(o.tap("simi_idx")) - 44600) < 100


def test_oracle_stage_failures_mirror_run_step():
    o = po.RosaOracle()
    assert not o.run(np.zeros((0, 3), np.float32)) and o.failed_stage == 1          # rosa.cpp:63-65
    assert not o.run(fx.cylinder(50)) and o.failed_stage == 1                       # < min_points (rosa.cpp:91-95)
    far = fx.cylinder(500) + np.array([100, 0, 0], np.float32)                      # beyond pts_dist_lim
    assert not o.run(far) and o.failed_stage == 1


def test_filter_is_stable_and_drops_nonfinite():
    P = fx.cylinder(2000, seed=5)
    P[10] = np.nan; P[20, 1] = np.inf; P[30] = [60, 0, 0]
    o = po.RosaOracle(stop_after_stage=1)
    assert o.run(P)
    idx = o.tap("filt_idx")
    assert len(idx) == 1997 and np.all(np.diff(idx) > 0) and not {10, 20, 30} & set(idx.tolist())
    assert_bit_equal(o.tap("filtered"), P[idx], "filtered")


def test_faithful_mode_is_statistically_close_to_canonical():
    """The reference's own unspecified orders (DFS sum order, libm trig/cbrt) are chaotic at ulp
    scale: report-level check only -- same M, same neighbour sets, |dV| small."""
    P = fx.cylinder()
    a = po.RosaOracle(); b = po.RosaOracle(faithful_normals=1, faithful_order=1, libm_cbrt=1)
    assert a.run(P) and b.run(P)
    assert len(a.tap("pts")) == len(b.tap("pts"))
    assert_bit_equal(a.tap("vox_keys"), b.tap("vox_keys"), "voxel keys")
    assert abs(len(a.tap("skelver")) - len(b.tap("skelver"))) <= 5
    va, vb = a.tap("skelver"), b.tap("skelver")
    d = np.sqrt(((va[:, None, :] - vb[None, :, :]) ** 2).sum(-1))
    assert max(d.min(1).max(), d.min(0).max()) < 2.0 * a.leaf_size      # Hausdorff within ~2 leaf


def test_graph_oracle_is_a_spanning_forest_with_canonical_edge_order():
    rng = np.random.default_rng(9)
    V = np.cumsum(rng.normal(size=(60, 3)) * 0.8 + np.array([0, 0, 1.5]), axis=0).astype(np.float32)
    g = po.GraphOracle(V, 3.0)
    e = g.tap("mst_edges").reshape(-1, 2); w = g.tap("mst_w")
    assert np.all(e[:, 0] < e[:, 1]) and np.all(np.diff(w) >= 0)
    # acyclic: union-find never sees an edge inside one component
    parent = list(range(60))
    def find(x):
        while parent[x] != x:
            x = parent[x]
        return x
    for u, v in e:
        ru, rv = find(u), find(v)
        assert ru != rv
        parent[rv] = ru
    t = g.tap("type"); off = g.tap("adj_off")
    deg = np.diff(off)
    assert np.array_equal(t, (deg == 1).astype(np.int32))        # gskel.cpp:554-562 fall-through: only leaves keep a type


def test_gskel_oracle_approves_on_third_update_and_respects_gnd_th():
    cfg = po.GSkelConfig.default(gnd_th=-1e9)
    g = po.GSkelOracle(cfg)
    pts = np.array([[0, 0, 0], [4, 0, 0], [8, 0, 0], [12, 0, 0], [16, 0, 0], [20, 0, 0]], np.float32)
    stages = []
    for _ in range(5):
        g.run(pts); stages.append(g.failed_stage)
    # tick 1 spawns, 2-3 update (trace P 3 -> .27 -> .14), tick 4 = 3rd update -> approved (trace .097 < .1)
    assert stages[:3] == [1, 1, 1]
    assert stages[3] == 4          # all vertices new -> vertex_merge aborts (>50 % new, gskel.cpp:286)
    assert len(g.tap("global")) == 6
    g2 = po.GSkelOracle(po.GSkelConfig.default())      # default gnd_th = 60 m rejects everything
    for _ in range(5):
        g2.run(pts)
    assert len(g2.tap("global")) == 0


def test_tf_points_matches_float64_rigid_transform():
    """orc_tf_points restates tf2::doTransform(PointCloud2) (gskel_node.cpp:137-138) in float32; against the float64
    rigid transform of the same quaternion it may differ by rounding only."""
    rng = np.random.default_rng(5)
    for f in (0, 17, 113, 240):
        P, R, t = fx.stream_frame(f, n=2000)
        q = fx.quat_xyzw_from_matrix(R)
        x, y, z, w = q
        Rq = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                       [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                       [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        assert np.abs(Rq - R).max() < 1e-12
        out = po.tf_points(P, t, q)
        ref = P.astype(np.float64) @ R.T + t
        assert out.dtype == np.float32 and np.abs(out - ref).max() < 2e-5
    # identity transform returns the input bits
    P = rng.normal(size=(100, 3)).astype(np.float32)
    assert np.array_equal(po.tf_points(P, [0, 0, 0], [0, 0, 0, 1]), P)
