"""GPU parity tests (run on a B200: python -m pytest tests -m gpu).  Every test calls the CUDA
path through the C ABI (libosep_skel.so via ctypes) and compares with the CPU oracle on the
same seeded input.  Bars: bit-exact for indices, keys, counts and connectivity; vertex
positions <= 1e-3 * leaf and orientations <= 1e-3 rad (BASELINE.json north_star) -- in
practice the floats are bit-identical too, and the tests assert that as well."""
import os

import numpy as np
import pytest

from conftest import assert_bit_equal

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

TAPS_EXACT = ["filt_idx", "filtered", "leaf_hist", "nvox_hist", "knn_full", "normals_full", "vox_keys", "vox_count", "pts", "nrms",
              "surf_idx", "simi_off", "simi_idx", "vset_iters", "active_size", "poor_idx", "pset_drosa", "pset", "seeds", "corresp",
              "skelver_sampled", "skelver"]


@pytest.fixture(scope="module")
def mods(built):
    import osep_path_planner_b200 as osep
    from oracle import pyoracle as po
    from osep_path_planner_b200 import fixtures as fx
    assert osep.load().osep_device_count() > 0, "no CUDA device: GPU tests need a B200"
    return osep, po, fx


def run_both(mods, P, cfgkw=None, debug=True):
    osep, po, fx = mods
    cfgkw = cfgkw or {}
    o = po.RosaOracle(po.RosaConfig.default(**cfgkw))
    ok_o = o.run(P)
    r = osep.Rosa(osep.RosaConfig(**cfgkw), capacity_points=max(1024, P.shape[0]))
    r.set_debug(debug)
    r.set_input(P)
    ok_g = r.rosa_run()
    return o, ok_o, r, ok_g


def check_all_taps(o, r):
    res = r.result
    assert res.n_filtered == len(o.tap("filt_idx")) and res.n_downsampled == len(o.tap("pts")) and res.n_vertices == len(o.tap("skelver"))
    assert np.float32(res.leaf_size) == np.float32(o.leaf_size) and np.float32(res.leaf_used) == np.float32(o.leaf_used)
    assert res.flags == o.flags
    for tap in TAPS_EXACT:
        a, b = r.tap(tap), o.tap(tap)
        if tap == "knn_full":
            a, b = a.reshape(-1), b.reshape(-1)
        assert_bit_equal(a, b, tap)
    # north_star tolerances, stated explicitly (they hold trivially when the bits match)
    leaf = o.leaf_size
    assert np.abs(r.tap("skelver") - o.tap("skelver")).max() <= 1e-3 * leaf
    va, vb = r.tap("vset"), o.tap("vset")
    cosang = np.clip(np.abs((va * vb).sum(1)), 0, 1)
    assert np.arccos(cosang).max() <= 1e-3
    # output_vertices(): V x 3 column-major (Eigen::MatrixXf layout, rosa.hpp:53)
    v = r.output_vertices()
    assert v.flags["F_CONTIGUOUS"] and v.dtype == np.float32
    assert_bit_equal(np.ascontiguousarray(v), o.tap("skelver"), "output_vertices")


@pytest.mark.parametrize("name", ["cyl5k", "cyl20k", "turbine100k"])
def test_full_pipeline_bit_exact(mods, name):
    osep, po, fx = mods
    P = {"cyl5k": lambda: fx.cylinder(5000, seed=3), "cyl20k": fx.cylinder, "turbine100k": fx.turbine}[name]()
    o, ok_o, r, ok_g = run_both(mods, P)
    assert ok_o and ok_g
    check_all_taps(o, r)
    r.close()


def test_golden_fixture_against_gpu(mods):
    osep, po, fx = mods
    g = np.load(os.path.join(GOLD, "cyl3k_k10_m300.npz"))
    r = osep.Rosa(osep.RosaConfig(max_points=300, ne_knn=10, nb_knn=6), capacity_points=4096)
    r.set_input(fx.cylinder(3000, seed=7))
    assert r.rosa_run()
    for tap in ["leaf_hist", "nvox_hist", "vox_keys", "vox_count", "pts", "nrms", "surf_idx", "vset", "pset", "poor_idx", "seeds", "corresp", "skelver"]:
        assert_bit_equal(r.tap(tap), g[tap], tap)
    assert_bit_equal(r.tap("simi_off"), g["simi_off"], "simi_off")
    assert_bit_equal(r.graph(3.0)["edges"].reshape(-1), g["mst_edges"], "mst_edges")
    r.close()


@pytest.mark.parametrize("cfgkw", [dict(ne_knn=10, nb_knn=6, max_points=500), dict(ne_knn=30, nb_knn=12, niter_drosa=3, niter_dcrosa=2, niter_smooth=1),
                                   dict(ne_knn=16, max_points=1500, radius_smooth=3.0, alpha_recenter=0.5),
                                   dict(niter_drosa=0, niter_dcrosa=0, niter_smooth=0), dict(niter_drosa=8, nb_knn=4, ne_knn=5),
                                   dict(niter_drosa=9, max_points=200)])
def test_non_default_configs(mods, cfgkw):
    osep, po, fx = mods
    o, ok_o, r, ok_g = run_both(mods, fx.turbine(30000, seed=11), cfgkw)
    assert ok_o and ok_g
    check_all_taps(o, r)
    r.close()


def test_stage_failures_mirror_run_step(mods):
    osep, po, fx = mods
    r = osep.Rosa(capacity_points=4096)
    r.input_cloud(0)
    assert not r.rosa_run() and r.result.status == 1                                   # empty cloud (rosa.cpp:63-65)
    r.set_input(fx.cylinder(50))
    assert not r.rosa_run() and r.result.status == 1 and r.result.n_filtered == 50     # < min_points (rosa.cpp:91-95)
    r.set_input(fx.cylinder(500) + np.array([100, 0, 0], np.float32))
    assert not r.rosa_run() and r.result.status == 1 and r.result.n_filtered == 0      # everything beyond pts_dist_lim
    assert r.output_vertices().shape == (0, 3)
    # the handle recovers on the next good frame
    P = fx.cylinder(3000, seed=1)
    r.set_input(P)
    assert r.rosa_run()
    o = po.RosaOracle(); assert o.run(P)
    assert_bit_equal(np.ascontiguousarray(r.output_vertices()), o.tap("skelver"), "vertices after failed frames")
    with pytest.raises(osep.OsepError):
        r.set_input(fx.cylinder(5000)); r.rosa_run()                                   # exceeds capacity: loud error
    r.close()


def test_ragged_input_nonfinite_far_points_duplicates_and_stride(mods):
    osep, po, fx = mods
    P = fx.cylinder(6000, seed=21)
    P[5] = np.nan; P[77, 2] = np.inf; P[100] = [0, 0, 80]; P[200:260] = P[300:360]     # NaN, inf, far, exact duplicates (kNN ties)
    P[1000:1100] = np.round(P[1000:1100] * 8) / 8                                       # lattice points -> equal distances
    o, ok_o, r, ok_g = run_both(mods, P)
    assert ok_o and ok_g
    check_all_taps(o, r)
    # 16-byte pcl::PointXYZ layout (stride 4) and packed xyz (stride 3) give identical results
    L = osep.load()
    import ctypes as C
    res = osep.RosaResult()
    P3 = np.ascontiguousarray(P)
    assert L.osep_rosa_run(r._h, P3.ctypes.data_as(C.c_void_p), P3.shape[0], 3, C.byref(res)) == 0
    assert_bit_equal(r.tap("skelver"), o.tap("skelver"), "stride 3")
    # wide host records (x, y, z, intensity, ring, time: 6 floats): only x, y, z travel to the device
    P6 = np.ascontiguousarray(np.concatenate([P, np.full((len(P), 3), 7.0, np.float32)], axis=1))
    assert L.osep_rosa_run(r._h, P6.ctypes.data_as(C.c_void_p), P6.shape[0], 6, C.byref(res)) == 0
    assert_bit_equal(r.tap("skelver"), o.tap("skelver"), "stride 6")
    r.close()


def test_minimum_size_cloud_and_tiny_k(mods):
    osep, po, fx = mods
    o, ok_o, r, ok_g = run_both(mods, fx.cylinder(120, seed=2, h=10.0), dict(max_points=40, min_points=100))
    assert ok_o == ok_g
    if ok_o:
        check_all_taps(o, r)
    else:
        assert r.result.status == o.failed_stage
    r.close()


def test_idempotent_and_handle_reuse(mods):
    osep, po, fx = mods
    r = osep.Rosa(capacity_points=120000)
    A, B = fx.turbine(), fx.cylinder()
    r.set_input(A); assert r.rosa_run(); va = np.ascontiguousarray(r.output_vertices()).copy()
    r.set_input(B); assert r.rosa_run(); vb = np.ascontiguousarray(r.output_vertices()).copy()
    r.set_input(A); assert r.rosa_run()
    assert_bit_equal(np.ascontiguousarray(r.output_vertices()), va, "frame A repeated after frame B")
    r.set_input(B); assert r.rosa_run()
    assert_bit_equal(np.ascontiguousarray(r.output_vertices()), vb, "frame B repeated")
    r.close()


def test_device_resident_entry_point_matches_host_entry_point(mods):
    import torch
    osep, po, fx = mods
    P = fx.to_xyz4(fx.cylinder())
    r = osep.Rosa(capacity_points=32768)
    r.set_input(P[:, :3]); assert r.rosa_run(); v_host = np.ascontiguousarray(r.output_vertices()).copy()
    t = torch.from_numpy(P).cuda()
    torch.cuda.synchronize()
    r.run_device(t.data_ptr(), P.shape[0], 4)
    assert r.finish()
    assert_bit_equal(np.ascontiguousarray(r.output_vertices()), v_host, "device entry point")
    assert r.result.gpu_ms > 0
    r.close()


def test_batched_frames_equal_single_frames(mods):
    osep, po, fx = mods
    frames = [fx.to_xyz4(fx.batch_frame(b, n=20000)) for b in range(6)] + [fx.to_xyz4(fx.cylinder(50))]
    r = osep.Rosa(capacity_points=32768)
    r.set_lanes(3)
    res, verts = r.run_batch(frames)
    assert [x.status for x in res[:6]] == [0] * 6 and res[6].status == 1
    for b in range(6):
        o = po.RosaOracle(); assert o.run(frames[b][:, :3].copy())
        assert_bit_equal(verts[b], o.tap("skelver"), "batch frame %d" % b)
        assert res[b].n_downsampled == len(o.tap("pts"))
    assert verts[6].shape == (0, 3)
    r.close()


def test_skeleton_graph_bit_exact(mods):
    osep, po, fx = mods
    rng = np.random.default_rng(5)
    for G in (1, 2, 9, 10, 11, 60, 333):
        V = np.cumsum(rng.normal(size=(G, 3)) * 0.8 + np.array([0, 0, 1.2]), axis=0).astype(np.float32)
        if G > 20:
            V[7] = V[3]                                        # duplicate vertex -> zero-length edge, distance ties
        g = osep.build_graph(V, 3.0)
        go = po.GraphOracle(V, 3.0)
        assert_bit_equal(g["edges"].reshape(-1), go.tap("mst_edges"), "mst edges G=%d" % G)
        assert_bit_equal(g["edge_w"], go.tap("mst_w"), "mst weights")
        assert_bit_equal(g["adj_off"], go.tap("adj_off"), "adj_off"); assert_bit_equal(g["adj_idx"], go.tap("adj_idx"), "adj_idx")
        assert_bit_equal(g["rng_off"], go.tap("rng_off"), "rng_off"); assert_bit_equal(g["rng_idx"], go.tap("rng_idx"), "rng_idx")
        assert_bit_equal(g["type"], go.tap("type"), "type")
        assert_bit_equal(g["joints"], go.tap("joints"), "joints"); assert_bit_equal(g["leafs"], go.tap("leafs"), "leafs")


def test_skeleton_graph_forests_ties_and_large_graphs(mods):
    """graph_adj (thread per vertex, register top-10) and mst (Boruvka rounds over the sorted keys) against the oracle's kd-tree
    + Kruskal on the cases a chain does not reach: a lattice (every weight tied: the (w,u,v) order decides), clusters farther
    apart than 2.5 x fuse_dist_th (a spanning FOREST), fewer vertices than k, non-finite vertices, and graphs whose keys and
    component arrays do not fit the shared-memory workspaces (global scratch path)."""
    osep, po, fx = mods
    rng = np.random.default_rng(11)
    gx, gy, gz = np.meshgrid(np.arange(6), np.arange(6), np.arange(6), indexing="ij")
    lattice = (np.stack([gx, gy, gz], -1).reshape(-1, 3) * 1.5).astype(np.float32)
    clusters = np.concatenate([rng.normal(size=(40, 3)) * 2.0 + c for c in ([0, 0, 0], [100, 0, 0], [0, 100, 50], [7.4, 0, 0])]).astype(np.float32)
    few = (rng.normal(size=(5, 3)) * 2.0).astype(np.float32)
    walk = np.cumsum(rng.normal(size=(2500, 3)) * 0.8 + np.array([0, 0, 1.2]), axis=0).astype(np.float32)
    blob = (rng.normal(size=(1500, 3)) * 6.0).astype(np.float32)
    for name, V in (("lattice", lattice), ("clusters", clusters), ("few", few), ("walk2500", walk), ("blob1500", blob)):
        g = osep.build_graph(V, 3.0)
        go = po.GraphOracle(V, 3.0)
        assert_bit_equal(g["rng_off"], go.tap("rng_off"), name + " rng_off"); assert_bit_equal(g["rng_idx"], go.tap("rng_idx"), name + " rng_idx")
        assert_bit_equal(g["edges"].reshape(-1), go.tap("mst_edges"), name + " mst edges")
        assert_bit_equal(g["edge_w"], go.tap("mst_w"), name + " mst weights")
        assert_bit_equal(g["adj_off"], go.tap("adj_off"), name + " adj_off"); assert_bit_equal(g["adj_idx"], go.tap("adj_idx"), name + " adj_idx")
        assert_bit_equal(g["type"], go.tap("type"), name + " type")
    n_comp = len(clusters) - len(osep.build_graph(clusters, 3.0)["edge_w"])
    assert n_comp >= 3, "the cluster case must be a forest (components: %d)" % n_comp


@pytest.mark.parametrize("mode", [dict(OSEP_DC_CLUSTER="8"), dict(OSEP_DC_CLUSTER="16"), dict(OSEP_VS_SEPARATE="1")])
def test_dcrosa_cluster_sizes_and_separate_vertex_sampling_agree(mods, monkeypatch, mode):
    """dcrosa + vertex_sampling as one cluster kernel of 8 CTAs (lanes of a batch), of 16 CTAs (single frames) and with
    vertex_sampling as its own five launches: every tap equals the oracle's in each mode."""
    osep, po, fx = mods
    for k, v in mode.items():
        monkeypatch.setenv(k, v)
    for P in (fx.turbine(60000, seed=9), fx.cylinder(15000, seed=3)):
        o, ok_o, r, ok_g = run_both(mods, P, debug=True)
        assert ok_o and ok_g
        check_all_taps(o, r)
        r.close()


def test_frame_to_graph(mods):
    osep, po, fx = mods
    P = fx.turbine(40000, seed=4)
    o, ok_o, r, ok_g = run_both(mods, P, debug=False)
    assert ok_o and ok_g
    g = r.graph(3.0); go = po.GraphOracle(o.tap("skelver"), 3.0)
    assert_bit_equal(g["edges"].reshape(-1), go.tap("mst_edges"), "frame graph edges")
    assert_bit_equal(g["type"], go.tap("type"), "frame graph types")
    r.close()


def test_in_frame_graph_equals_separate_graph_call(mods):
    """osep_rosa_set_graph: graph_adj + mst inside the frame's launch sequence (vertex count read on the device) give the same
    views as the separate osep_rosa_graph call and as the oracle, across frames with different vertex counts, graph replays
    included; a different fuse_dist_th falls back to the separate call."""
    osep, po, fx = mods
    r = osep.Rosa(capacity_points=40000)
    r.set_graph(True, 3.0)
    for k, P in enumerate([fx.cylinder(12000, seed=2), fx.turbine(40000, seed=3), fx.turbine(25000, seed=4), fx.cylinder(9000, seed=5), fx.turbine(33000, seed=6)]):
        o = po.RosaOracle(); assert o.run(P)
        r.set_input(P); assert r.rosa_run()
        g = r.graph(3.0)
        go = po.GraphOracle(o.tap("skelver"), 3.0)
        assert g["n_vertices"] == len(o.tap("skelver"))
        for key, tap in [("edges", "mst_edges"), ("adj_idx", "adj_idx"), ("rng_idx", "rng_idx"), ("type", "type"), ("joints", "joints"), ("leafs", "leafs")]:
            assert_bit_equal(np.asarray(g[key]).reshape(-1), go.tap(tap), "%s frame %d" % (key, k))
        g2 = r.graph(2.0); go2 = po.GraphOracle(o.tap("skelver"), 2.0)
        assert_bit_equal(np.asarray(g2["edges"]).reshape(-1), go2.tap("mst_edges"), "fuse 2.0 frame %d" % k)
    r.close()


def test_gskel_stateful_replay(mods):
    """Stream of frames (BASELINE configs[2] in miniature): Rosa vertices -> analytic sensor->odom
    transform -> GSkel::gskel_run, oracle and CUDA library side by side, compared every tick."""
    osep, po, fx = mods
    from osep_path_planner_b200 import GSkel, GSkelConfig
    cfg_kw = dict(gnd_th=-100.0)           # fixtures sit below the default 60 m ground threshold
    gs = GSkel(GSkelConfig(**cfg_kw)); go = po.GSkelOracle(po.GSkelConfig.default(**cfg_kw))
    r = osep.Rosa(capacity_points=32768)
    published = 0
    for f in range(0, 24):
        P, R, t = fx.stream_frame(f * 3, n=20000)
        r.set_input(P)
        if not r.rosa_run():
            continue
        v = np.ascontiguousarray(r.output_vertices())
        v_odom = (v.astype(np.float64) @ R.T + t).astype(np.float32)
        ok_g = gs.gskel_run() if gs.set_input(v_odom) is not None else False
        ok_o = go.run(v_odom)
        assert ok_g == ok_o and gs.failed_stage == go.failed_stage, "tick %d: stage %d vs %d" % (f, gs.failed_stage, go.failed_stage)
        assert_bit_equal(gs.output_gskel()[:, :3].copy(), go.tap("global"), "global vertices tick %d" % f)
        assert_bit_equal(gs.tap("prelim"), go.tap("prelim"), "prelim vertices tick %d" % f)
        for tap in ["adj_off", "adj_idx", "joints", "leafs", "new_idx", "type", "obs", "branch_off", "branch_idx"]:
            assert_bit_equal(gs.tap(tap), go.tap(tap), "%s tick %d" % (tap, f))
        published += int(ok_g)
    assert len(go.tap("global")) > 5
    gs.close(); r.close()


def test_full_size_properties_turbine100k(mods):
    """Size-independent properties at BASELINE.json's full size (100k points)."""
    osep, po, fx = mods
    P = fx.turbine()
    r = osep.Rosa(capacity_points=100000); r.set_debug(True); r.set_input(P); assert r.rosa_run()
    idx = r.tap("filt_idx"); assert np.all(np.diff(idx) > 0)                             # stable compaction
    keys = r.tap("vox_keys"); assert np.all(np.diff(keys.astype(np.int64)) > 0)          # ascending distinct voxel keys
    assert r.tap("vox_count").sum() == len(idx)                                          # every point lands in one voxel
    knn = r.tap("knn_full").reshape(len(idx), -1)
    assert np.array_equal(knn[:, 0], np.arange(len(idx)))                                # nearest neighbour of a point is itself
    F = r.tap("filtered")
    d = ((F[knn[::97]] - F[::97, None, :]) ** 2).sum(-1); assert np.all(np.diff(d, axis=1) >= 0)   # ascending distances
    n = r.tap("normals_full"); assert np.abs(np.linalg.norm(n, axis=1) - 1).max() < 1e-3
    assert np.all((n * (-F)).sum(1) >= 0)                                                # flipped towards the viewpoint (origin)
    corresp = r.tap("corresp"); assert corresp.min() >= 0 and corresp.max() == r.result.n_vertices - 1
    seeds = r.tap("seeds"); assert np.all(np.diff(seeds) > 0)                            # greedy cover picks ascending seeds
    pc = r.tap("pset"); R2 = np.float32(r.result.leaf_size) ** 2
    dd = ((pc[seeds][:, None, :] - pc[seeds][None, :, :]) ** 2).sum(-1) + np.eye(len(seeds)) * 1e9
    assert dd.min() >= R2 * 0.999                                                        # no seed lies inside another seed's cover
    r.close()


def test_cpp_shim_demo_node_runs(mods):
    """The C++ look-alike classes (host/rosa.hpp, host/gskel.hpp) drive the library exactly like the
    reference's process_tick functions and print the reference's stdout protocol."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "osep_path_planner_b200", "host", "demo_node")
    out = subprocess.run([exe, "20000", "5"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.splitlines()
    assert "PointCloud size before downsampling: 20000" in lines
    for name in ["preprocess", "rosa_init", "similarity_neighbor_extraction", "drosa", "dcrosa", "vertex_sampling", "vertex_smooth"]:
        assert "[SUCCESS] " + name in lines
    assert any(l.startswith("Number of vertices: ") for l in lines) and any(l.startswith("[ROSA] Time Elapsed:") for l in lines)
    assert any(l.startswith("[FAILED] increment_skeleton") or l.startswith("[SUCCESS] increment_skeleton") for l in lines)
    assert any(l.startswith("global skeleton:") or "[FAILED] vertex_merge" in l for l in lines)


def test_tiled_map_per_tile_parity(mods):
    """configs[4] in miniature on one GPU: a two-structure map split into 2 slabs + halo; every tile (tile + halo
    input) must match the oracle bit-for-bit, and the seam graph over the union of owned vertices too."""
    osep, po, fx = mods
    from osep_path_planner_b200 import tiled
    A = fx.turbine(40000, seed=3)
    B = fx.cylinder(20000, seed=4, cx=15.0) + np.array([0, 40, 0], np.float32)
    P = np.concatenate([A, B]).astype(np.float32)
    axis, bounds = tiled.split_axis_and_bounds(P, 2)
    halo = tiled.default_halo(P)
    cfg_kw = dict(pts_dist_lim=1000.0)
    r = osep.Rosa(osep.RosaConfig(**cfg_kw), capacity_points=len(P))
    owned = []
    for t in range(2):
        own = tiled.slab_of(P, axis, bounds, t)
        c = P[:, axis]
        nb = P[((c >= bounds[t] - halo) & (c < bounds[t])) | ((c >= bounds[t + 1]) & (c < bounds[t + 1] + halo))]
        tile = np.ascontiguousarray(np.concatenate([own, nb]))
        r.set_input(tile); ok = r.rosa_run()
        o = po.RosaOracle(po.RosaConfig.default(**cfg_kw)); ok_o = o.run(tile)
        assert ok and ok_o
        assert_bit_equal(np.ascontiguousarray(r.output_vertices()), o.tap("skelver"), "tile %d vertices" % t)
        owned.append(tiled.owned_filter(o.tap("skelver"), axis, bounds[t], bounds[t + 1]))
    allv = np.ascontiguousarray(np.concatenate(owned))
    assert len(allv) > 20
    g = osep.build_graph(allv, 3.0); go = po.GraphOracle(allv, 3.0)
    assert_bit_equal(g["edges"].reshape(-1), go.tap("mst_edges"), "seam graph edges")
    r.close()


def test_device_resident_tile_primitives(mods):
    """The C-ABI pieces of the device-resident tiled path (configs[4]): stable range compaction, the leaf-only pass, the
    device copy of the vertices, the device graph build and TileRunner on one rank -- each against numpy / the oracle."""
    import ctypes as C
    import torch
    osep, po, fx = mods
    from osep_path_planner_b200 import tiled
    P = fx.turbine(50000, seed=12)
    t = torch.from_numpy(P).cuda()
    for axis, lo, hi in ((0, -1.0, 2.5), (2, 10.0, 1e30), (1, -1e30, 0.0)):
        sel = tiled.select_range(t, axis, lo, hi).cpu().numpy()
        ref = P[(P[:, axis] >= lo) & (P[:, axis] < hi)]
        assert_bit_equal(sel, ref, "select_range axis %d" % axis)            # same rows, same order
    sel4 = tiled.select_range(t, 0, 0.0, 1.0, out_stride=4).cpu().numpy()
    assert sel4.shape[1] == 4 and np.all(sel4[:, 3] == 1.0)
    # leaf-only pass == the oracle's converged leaf on the same cloud
    cfg = dict(pts_dist_lim=1.0e4)
    o = po.RosaOracle(po.RosaConfig.default(**cfg)); assert o.run(P)
    run = tiled.TileRunner(0, osep.RosaConfig(**cfg))
    assert run.own_leaf(t) == np.float32(o.leaf_size)
    # one rank: the tile is the slab, every vertex is owned; device vertices == host vertices == oracle
    res = run.run(t, 2, -np.inf, np.inf)
    assert res["ok"] and res["halo_points"] == 0 and res["tile_points"] == len(P)
    v_dev = res["vertices"].cpu().numpy()
    assert_bit_equal(v_dev[:, :3], o.tap("skelver"), "device vertices")
    assert np.all(v_dev[:, 3] == 1.0)
    go = po.GraphOracle(o.tap("skelver"), 3.0)
    assert_bit_equal(res["graph"]["edges"].reshape(-1), go.tap("mst_edges"), "graph over device vertices")
    # owned mask: a sub-interval keeps exactly the vertices inside it, in order
    sk = o.tap("skelver"); mid = float(np.median(sk[:, 2]))
    res2 = run.run(t, 2, -np.inf, mid, halo=1.0)
    assert_bit_equal(res2["local"].cpu().numpy()[:, :3], sk[sk[:, 2] < mid], "owned vertices")
    run.close()


def test_map_accumulation_and_device_tf(mods):
    """the TF kernel, the device hand-off Rosa -> TF -> GSkel and the accumulated voxel map against their
    CPU restatements (tests/pyref/map_ref.py, the oracle's orc_tf_points and the host entry point osep_gskel_run_tf)."""
    import ctypes as C
    import torch
    osep, po, fx = mods
    from osep_path_planner_b200.map import SkelMap
    from osep_path_planner_b200 import GSkel, GSkelConfig
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "pyref"))
    import map_ref
    L = osep.load()
    # (1) TF kernel == float32 restatement, bit for bit
    P, Rm, tv = fx.stream_frame(40, n=20000)
    q = np.array([0.1, -0.2, 0.3, 0.9]); q /= np.linalg.norm(q)
    t = np.array([12.5, -3.25, 40.0])
    tin = torch.from_numpy(np.ascontiguousarray(P)).cuda(); tout = torch.empty((len(P), 3), dtype=torch.float32, device="cuda")
    rc = L.osep_tf_points_device(0, C.c_void_p(tin.data_ptr()), len(P), 3, t.ctypes.data_as(C.c_void_p), q.ctypes.data_as(C.c_void_p),
                                 C.c_void_p(tout.data_ptr()), 3, None)
    assert rc == 0
    torch.cuda.synchronize()
    assert_bit_equal(tout.cpu().numpy(), map_ref.tf_points(P, t, q), "tf kernel")
    # (2) accumulated map: 6 overlapping stream frames, each with its pose; same voxels, same points, same order
    m = SkelMap(capacity_points=400000, max_frame_points=30000, voxel_size=0.25)
    ref = map_ref.MapRef(0.25)
    for f in range(6):
        Pf, Rm, tv = fx.stream_frame(f * 20, n=25000)
        qf = np.array([0.0, 0.0, np.sin(0.05 * f), np.cos(0.05 * f)]); tf_ = np.array([0.3 * f, -0.1 * f, 0.0])
        if f == 3:
            Pf = Pf.copy(); Pf[::997] = np.nan                                   # non-finite points are dropped
        got = m.insert(Pf, tf_, qf) if f % 2 == 0 else m.insert_device(torch.from_numpy(np.ascontiguousarray(Pf)).cuda().data_ptr(), len(Pf), 3, tf_, qf)
        assert got == ref.insert(Pf, tf_, qf), "new voxels of frame %d" % f
    mp = m.points_tensor().cpu().numpy()
    assert_bit_equal(mp[:, :3], ref.array(), "accumulated map points (order = first appearance)")
    assert np.all(mp[:, 3] == 1.0) and len(mp) > 20000
    # the map feeds the tiled path without leaving the device
    from osep_path_planner_b200 import tiled
    run = tiled.TileRunner(0, osep.RosaConfig(pts_dist_lim=1.0e4))
    res = run.run(m.points_tensor(), 0, -np.inf, np.inf)
    o = po.RosaOracle(po.RosaConfig.default(pts_dist_lim=1.0e4)); ok_o = o.run(ref.array())
    assert res["ok"] == ok_o
    if ok_o:
        assert_bit_equal(res["vertices"].cpu().numpy()[:, :3], o.tap("skelver"), "skeleton of the accumulated map")
    run.close(); m.close()
    # (3) Rosa -> TF -> GSkel on the device == host TF entry point, tick by tick
    r = osep.Rosa(capacity_points=60000)
    ga, gb = GSkel(GSkelConfig(gnd_th=-100.0)), GSkel(GSkelConfig(gnd_th=-100.0))
    for f in range(8):
        Pf, Rm, tv = fx.stream_frame(f * 10, n=30000)
        r.set_input(Pf)
        if not r.rosa_run():
            continue
        qf = np.array([0.0, 0.0, np.sin(0.02 * f), np.cos(0.02 * f)]); tf_ = np.array([0.5 * f, 0.2 * f, 1.0])
        ga.set_input(np.ascontiguousarray(r.output_vertices())); oka = ga.gskel_run_tf(tf_, qf)
        okb = gb.gskel_run_tf_device(r, tf_, qf)
        assert bool(okb) == bool(oka)
        assert_bit_equal(np.ascontiguousarray(ga.output_gskel()), np.ascontiguousarray(gb.output_gskel()), "global skeleton tick %d" % f)
        ea, eb = ga.graph(), gb.graph()
        assert_bit_equal(ea["edges"].reshape(-1), eb["edges"].reshape(-1), "global graph tick %d" % f)
    ga.close(); gb.close(); r.close()


def test_large_voxels_global_sort_path(mods):
    """max_points small relative to N => voxels with more than 2048 points (the in-place global bitonic path)."""
    osep, po, fx = mods
    o, ok_o, r, ok_g = run_both(mods, fx.cylinder(60000, seed=8), dict(max_points=20, min_points=10), debug=False)
    assert ok_o == ok_g and r.result.status == o.failed_stage
    assert r.tap("vox_count").max() > 2048
    for tap in ["vox_keys", "vox_count", "pts", "nrms"]:
        assert_bit_equal(r.tap(tap), o.tap(tap), tap)
    if ok_o:
        assert_bit_equal(r.tap("skelver"), o.tap("skelver"), "skelver")
    r.close()


def test_many_frames_bit_exact(mods):
    """Breadth: different viewpoints / stand-offs / sizes (BASELINE configs[3]-style frames and stream frames)."""
    osep, po, fx = mods
    r = osep.Rosa(capacity_points=60000)
    frames = [fx.batch_frame(b, n=30000) for b in range(8)] + [fx.stream_frame(f, n=25000)[0] for f in (0, 60, 125, 190)]
    for k, P in enumerate(frames):
        r.set_input(P); ok = r.rosa_run()
        o = po.RosaOracle(); ok_o = o.run(P)
        assert ok == ok_o and r.result.status == o.failed_stage, "frame %d" % k
        assert r.result.flags == o.flags and r.result.n_downsampled == len(o.tap("pts"))
        assert_bit_equal(np.ascontiguousarray(r.output_vertices()), o.tap("skelver"), "frame %d vertices" % k)
        assert_bit_equal(r.tap("pset"), o.tap("pset"), "frame %d pset" % k)
        assert_bit_equal(r.tap("vset"), o.tap("vset"), "frame %d vset" % k)
    r.close()


def test_fused_and_unfused_drosa_agree(mods):
    """The persistent cooperative drosa kernel and the per-iteration fallback kernels give identical bits."""
    osep, po, fx = mods
    P = fx.turbine(50000, seed=9)
    outs = []
    for unfused in ("0", "1"):
        os.environ["OSEP_DROSA_UNFUSED"] = unfused
        try:
            r = osep.Rosa(capacity_points=50000); r.set_debug(True)
            r.set_input(P); assert r.rosa_run()
            outs.append((r.tap("vset_iters"), r.tap("pset_drosa"), np.ascontiguousarray(r.output_vertices()).copy()))
            r.close()
        finally:
            os.environ.pop("OSEP_DROSA_UNFUSED", None)
    for a, b, what in zip(outs[0], outs[1], ["vset_iters", "pset_drosa", "vertices"]):
        assert_bit_equal(a, b, what)
    o = po.RosaOracle(); assert o.run(P)
    assert_bit_equal(outs[0][2], o.tap("skelver"), "vertices vs oracle")


@pytest.mark.parametrize("case", ["forced_small_budget", "stream840_unconverged", "stream840_unconverged_unfused"])
def test_bit_matrix_in_global_memory_path_is_bit_exact(mods, monkeypatch, case):
    """Frames whose leaf loop does not converge keep M ~ 1200-1300 points: the M x M/32 similarity bit matrix then no longer
    fits the seed warps' shared memory and the active-set search reads its rows from global memory (list-driven, 16-byte
    row loads).  Forced on ordinary clouds by a small shared-memory budget, and reached naturally by a frame of the 50k-point
    stream (fused kernel and per-iteration kernels); every tap equals the oracle's."""
    osep, po, fx = mods
    if case.startswith("forced"):
        monkeypatch.setenv("OSEP_FUSED_SMEM_KB", "100")
        clouds = [fx.turbine(60000, seed=9), fx.cylinder(5000, seed=3)]
    else:
        if case.endswith("unfused"):
            monkeypatch.setenv("OSEP_DROSA_UNFUSED", "1")
        clouds = [fx.stream_frame(840, n=50000, n_frames=1000)[0]]
    for P in clouds:
        o, ok_o, r, ok_g = run_both(mods, P, debug=True)
        assert ok_o and ok_g
        if not case.startswith("forced"):
            assert r.result.n_downsampled > 1150, "this frame is here because its bit matrix does not fit (M = %d)" % r.result.n_downsampled
        check_all_taps(o, r)
        r.close()


def test_sparse_far_points_take_the_exact_knn_path_bit_exact(mods):
    """A frame whose last part is cut short (a stream frame truncated inside its third blade: 363 points spread over 28 m)
    has queries whose 20 nearest neighbours lie many grid cells away: they reach k_knn_exact (warp per query, ring shells,
    brute force beyond).  Neighbour lists, normals and everything downstream equal the oracle's; the frame stays fast."""
    osep, po, fx = mods
    P = fx.stream_frame(375, n=50000)[0][:42029]
    o, ok_o, r, ok_g = run_both(mods, P, debug=True)
    assert ok_o and ok_g
    assert int(r.tap("knn_stats")[1]) > 100, "expected the exact path to be exercised"
    check_all_taps(o, r)
    r.close()
    r = osep.Rosa(capacity_points=50000); r.set_input(P)
    for _ in range(4): r.rosa_run()
    assert r.result.gpu_ms < 2.0, "sparse points made the frame take %.2f ms" % r.result.gpu_ms
    r.close()


def _weird_clouds(fx):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from pyref import scenarios
    return scenarios.weird_clouds(fx)


@pytest.mark.parametrize("name", ["outliers", "plane", "line", "two_clusters", "duplicates", "lattice"])
def test_degenerate_geometries(mods, name):
    """Shapes the reference never meets on a turbine: the CUDA path must agree with the oracle on the stage that
    fails (RUN_STEP semantics) or on every bit of the result -- or fail loudly with a capacity error, never silently."""
    osep, po, fx = mods
    P = _weird_clouds(fx)[name]
    o = po.RosaOracle(); ok_o = o.run(P)
    r = osep.Rosa(capacity_points=len(P) + 64); r.set_debug(True)
    r.set_input(P)
    ok_g = r.rosa_run()                       # none of these clouds may hit a capacity limit (a limit would raise OsepError -3 here)
    st = r.result.status
    assert st >= 0, "%s: device-side error %d: %s" % (name, st, osep.last_error() if hasattr(osep, "last_error") else "")
    assert ok_g == ok_o and st == o.failed_stage, "%s: gpu status %d vs oracle stage %d" % (name, st, o.failed_stage)
    for tap in ["filt_idx", "leaf_hist", "nvox_hist", "knn_full", "normals_full", "vox_keys", "pts", "nrms"]:
        a, b = r.tap(tap), o.tap(tap)
        if tap == "knn_full":
            a, b = a.reshape(-1), b.reshape(-1)
        assert_bit_equal(a, b, "%s %s" % (name, tap))
    if ok_o:
        assert_bit_equal(r.tap("simi_idx"), o.tap("simi_idx"), name + " simi")
        assert_bit_equal(r.tap("pset"), o.tap("pset"), name + " pset")
        assert_bit_equal(np.ascontiguousarray(r.output_vertices()), o.tap("skelver"), name + " vertices")
        assert r.result.flags == o.flags
    r.close()


def test_frame_graph_replay_equals_plain_launches(mods):
    """One frame = one CUDA graph whose shape does not depend on the frame: n, stride and the buffer address travel through a
    device-side FrameParams block refreshed by a copy node (api.cu rosa_enqueue).  20 frames with 20 different point counts
    (a LiDAR stream never repeats n), different buffers and strides: all but the first two are pure graph replays, and every
    one matches the oracle bit-for-bit."""
    osep, po, fx = mods
    r = osep.Rosa(capacity_points=32000)
    rng = np.random.default_rng(5)
    sizes = [int(x) for x in rng.integers(9000, 32000, 20)]
    assert len(set(sizes)) == 20
    for k, n in enumerate(sizes):
        P = fx.turbine(n, seed=100 + k) if k % 3 else fx.cylinder(n, seed=100 + k)
        o = po.RosaOracle(); assert o.run(P)
        r.set_input(P if k % 2 else fx.to_xyz4(P))                 # stride 3 and stride 4 alternate
        assert r.rosa_run()
        assert_bit_equal(np.ascontiguousarray(r.output_vertices()), o.tap("skelver"), "frame %d (n=%d)" % (k, n))
    assert int(r.tap("graph_replays")[0]) >= 18, "frames with different n must replay the same graph"
    r.set_debug(True)                                              # debug mode = plain launches: same results
    P = fx.turbine(sizes[3], seed=103)
    o = po.RosaOracle(); assert o.run(P)
    r.set_input(P); assert r.rosa_run()
    assert_bit_equal(np.ascontiguousarray(r.output_vertices()), o.tap("skelver"), "plain launches")
    r.close()


def test_pointcloud2_ingest_equals_packed_input(mods):
    """osep_rosa_run_pointcloud2: raw PointCloud2 records (x,y,z at an offset inside a wider
    point_step, e.g. x y z intensity ring time) give the same frame as the packed pcl::PointXYZ buffer."""
    osep, po, fx = mods
    P = fx.turbine(25000, seed=21)
    o = po.RosaOracle(); assert o.run(P)
    r = osep.Rosa(capacity_points=25000)
    for point_step, off_x in ((16, 0), (32, 0), (48, 8), (12, 0)):
        rec = np.zeros((P.shape[0], point_step), np.uint8)
        rec[:] = 0xAB                                              # garbage in the other fields (incl. NaN-looking bytes)
        rec[:, off_x:off_x + 12] = P.astype(np.float32).view(np.uint8).reshape(-1, 12)
        assert r.rosa_run_pointcloud2(rec, P.shape[0], point_step, off_x)
        assert_bit_equal(np.ascontiguousarray(r.output_vertices()), o.tap("skelver"), "point_step %d" % point_step)
    L = osep.load(); import ctypes as C
    res = osep.RosaResult(); rec = np.zeros(64, np.uint8)
    assert L.osep_rosa_run_pointcloud2(r._h, rec.ctypes.data_as(C.c_void_p), 2, 18, 0, 4, 8, C.byref(res)) == -2      # point_step % 4
    assert L.osep_rosa_run_pointcloud2(r._h, rec.ctypes.data_as(C.c_void_p), 2, 16, 0, 8, 4, C.byref(res)) == -2      # x,y,z not consecutive
    r.close()


def test_gskel_run_tf_equals_transform_then_run(mods):
    """osep_gskel_run_tf = tf2::doTransform arithmetic + gskel_run, against the oracle's restatement."""
    osep, po, fx = mods
    from osep_path_planner_b200 import GSkel, GSkelConfig
    cfg_kw = dict(gnd_th=-100.0)
    gs = GSkel(GSkelConfig(**cfg_kw)); go = po.GSkelOracle(po.GSkelConfig.default(**cfg_kw))
    r = osep.Rosa(capacity_points=32768)
    for f in range(0, 12):
        P, R, t = fx.stream_frame(f * 3, n=20000)
        q = fx.quat_xyzw_from_matrix(R)
        r.set_input(P)
        if not r.rosa_run():
            continue
        v = np.ascontiguousarray(r.output_vertices())
        gs.set_input(v)
        ok_g = gs.gskel_run_tf(t, q)
        ok_o = go.run(po.tf_points(v, t, q))
        assert ok_g == ok_o and gs.failed_stage == go.failed_stage
        assert_bit_equal(gs.output_gskel()[:, :3].copy(), go.tap("global"), "global vertices tick %d" % f)
        assert_bit_equal(gs.tap("prelim"), go.tap("prelim"), "prelim vertices tick %d" % f)
    assert len(go.tap("prelim")) > 5
    gs.close(); r.close()


def test_knn_overflow_and_exact_paths_are_bit_exact(mods, monkeypatch):
    """The kNN stage has a fast path and fall-backs.  Thread-per-query scan (K <= 24): one attempt over the 27-cell block with
    the cell's pilot radius, k_knn_big (warp per query, 27 then 125 cells) for the misses, then k_knn_exact (ring search).  Warp-per-cell kernels
    (OSEP_KNN_WARP=1, and K > 24): k_knn_cells (<= 512 staged candidates), k_knn_big, k_knn_exact.  Deliberately fine and
    coarse grids (OSEP_KNN_PPC) push most queries through the fall-backs; neighbour lists and normals stay bit-identical
    to the oracle."""
    osep, po, fx = mods
    P = fx.turbine(60000, seed=31)
    o = po.RosaOracle(); assert o.run(P)
    seen = np.zeros(5, np.int64)
    for warp, ppc in (("0", "1"), ("0", "2"), ("0", "40"), ("0", "400"), ("1", "40"), ("1", "400")):
        monkeypatch.setenv("OSEP_KNN_PPC", ppc); monkeypatch.setenv("OSEP_KNN_WARP", warp)
        r = osep.Rosa(capacity_points=60000); r.set_debug(True); r.set_input(P)
        assert r.rosa_run()
        st = r.tap("knn_stats")
        if warp == "1":
            assert st[2] > 0 or st[1] > 0, "the overflow paths were not exercised: %s" % (st,)
        else:
            seen[:len(st)] += np.asarray(st, np.int64)
        assert_bit_equal(r.tap("knn_full").reshape(-1), o.tap("knn_full").reshape(-1), "knn_full ppc=%s warp=%s" % (ppc, warp))
        assert_bit_equal(r.tap("normals_full"), o.tap("normals_full"), "normals ppc=" + ppc)
        assert_bit_equal(np.ascontiguousarray(r.output_vertices()), o.tap("skelver"), "vertices ppc=" + ppc)
        r.close()
    assert seen[1] > 0 and seen[4] > 0, "scan fall-backs (warp-per-query, exact) were not all exercised: %s" % (seen,)


@pytest.mark.parametrize("cfgkw,cloud", [
    (dict(max_points=1500), ("turbine", 60000, 3)),            # M ~ 1560: bit matrix no longer fits the worker CTAs' shared memory
    (dict(max_points=1536), ("cylinder", 40000, 4)),           # largest max_points the handle accepts
    (dict(max_points=120, min_points=50), ("cylinder", 3000, 5)),
    (dict(niter_drosa=8), ("cylinder", 20000, 7)),             # MAX_FUSED_ITERS chains
    (dict(niter_drosa=1, niter_dcrosa=1, niter_smooth=0), ("cylinder", 20000, 8)),
    (dict(nb_knn=20, ne_knn=40), ("turbine", 30000, 9)),       # generic (non-10) surf lists in the chain walker, K = 64 kNN variant
])
def test_configuration_edges_production_path(mods, cfgkw, cloud):
    """Edges of the configuration space through the production path (CUDA graph replay on the third run, no debug taps)."""
    osep, po, fx = mods
    kind, n, seed = cloud
    P = fx.turbine(n, seed=seed) if kind == "turbine" else fx.cylinder(n, seed=seed)
    o = po.RosaOracle(po.RosaConfig.default(**cfgkw)); ok_o = o.run(P)
    r = osep.Rosa(osep.RosaConfig(**cfgkw), capacity_points=n)
    r.set_input(P)
    for _ in range(3):
        ok_g = r.rosa_run()
        assert ok_g == ok_o
        assert_bit_equal(np.ascontiguousarray(r.output_vertices()), o.tap("skelver"), "vertices %s" % (cfgkw,))
    assert r.result.n_downsampled == len(o.tap("pts")) and r.result.n_simi == len(o.tap("simi_idx"))
    r.close()


# ---------------------------------------------------------------------------------------------------------------
# The library against tests/pyref -- the independently written numpy / Python restatement of the reference (not the C++
# oracle): discrete Rosa stages and the full stateful GSkel replay.
def _pyref():
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from pyref import gskel_ref, rosa_ref, scenarios
    return gskel_ref, rosa_ref, scenarios


def _pyref_float():
    _pyref()
    from pyref import rosa_float_ref
    return rosa_float_ref


@pytest.mark.parametrize("name", ["cyl20k", "turbine100k", "stream131"])
def test_library_matches_pyref_on_every_rosa_stage(mods, name):
    """the CUDA library against tests/pyref directly (not through the C++ oracle): the discrete stages (rosa_ref) and the
    float stages up to the skeleton vertices (rosa_float_ref: orientations after every drosa iteration, active-set sizes,
    positions after drosa / dcrosa, greedy cover, vertices before / after vertex_smooth), bit for bit"""
    osep, po, fx = mods
    _, rr, _ = _pyref()
    rf = _pyref_float()
    P = {"cyl20k": fx.cylinder, "turbine100k": fx.turbine, "stream131": lambda: fx.stream_frame(131, n=30000)[0]}[name]()
    ref = rr.RosaRef(); assert ref.run(P)
    r = osep.Rosa(capacity_points=P.shape[0]); r.set_debug(True); r.set_input(P); assert r.rosa_run()
    assert np.float32(r.result.leaf_size) == ref.leaf_size and np.float32(r.result.leaf_used) == ref.leaf_used
    for tap in ["filt_idx", "filtered", "knn_full", "normals_full", "leaf_hist", "nvox_hist", "vox_keys", "vox_count", "pts", "nrms",
                "surf_idx", "simi_off", "simi_idx"]:
        a, b = r.tap(tap), ref.t[tap]
        assert_bit_equal(a.reshape(-1), b.reshape(-1), "%s vs pyref on %s" % (tap, name))
    ft = rf.RosaFloatRef(ref.t["pts"], ref.t["nrms"], ref.t["surf_idx"], ref.t["simi_off"], ref.t["simi_idx"], ref.leaf_size).run()
    for tap in ["vset_iters", "active_size", "poor_idx", "pset_drosa", "pset", "seeds", "corresp", "skelver_sampled", "skelver"]:
        assert_bit_equal(r.tap(tap).reshape(-1), np.asarray(ft[tap]).reshape(-1), "%s vs pyref on %s" % (tap, name))
    assert_bit_equal(np.ascontiguousarray(r.output_vertices()).reshape(-1), ft["skelver"].reshape(-1), "output_vertices vs pyref on %s" % name)
    # ... and on to the skeleton graph (graph_adj + mst + graph_decomp, gskel.cpp:201-281,539-567) of those vertices
    gr_, _, _ = _pyref()
    g, ref_g = r.graph(3.0), gr_.skeleton_graph(ft["skelver"], 3.0)
    assert_bit_equal(g["edges"].reshape(-1), ref_g["edges"].reshape(-1), "mst edges vs pyref on %s" % name)
    assert_bit_equal(g["edge_w"], ref_g["edge_w"], "mst weights vs pyref on %s" % name)
    assert_bit_equal(g["type"], ref_g["type"], "vertex types vs pyref on %s" % name)
    assert [list(g["rng_idx"][g["rng_off"][i]:g["rng_off"][i + 1]]) for i in range(len(ref_g["rng"]))] == ref_g["rng"], "rng lists vs pyref"
    r.close()


def _compare_gskel_with_pyref(gs, ref, what):
    assert_bit_equal(gs.output_gskel()[:, :3].copy(), ref.positions(), "global vertices " + what)
    assert_bit_equal(gs.tap("prelim"), ref.prelim_positions(), "prelim vertices " + what)
    off, idx = ref.csr(ref.adj)
    assert_bit_equal(gs.tap("adj_off"), off, "adj_off " + what); assert_bit_equal(gs.tap("adj_idx"), idx, "adj_idx " + what)
    off, idx = ref.csr(ref.branches)
    assert_bit_equal(gs.tap("branch_off"), off, "branch_off " + what); assert_bit_equal(gs.tap("branch_idx"), idx, "branch_idx " + what)
    for tap, val in [("joints", ref.joints), ("leafs", ref.leafs), ("new_idx", ref.new_idx), ("type", [v.type for v in ref.glob]),
                     ("obs", [v.obs_count for v in ref.glob]), ("prelim_frozen", [int(v.frozen) for v in ref.prelim])]:
        assert_bit_equal(gs.tap(tap), np.array(val, np.int32), tap + " " + what)


def test_gskel_library_matches_pyref_over_70_ticks_with_merges_and_prunes(mods):
    """G1-G10 against the independent Python state machine (tests/pyref/gskel_ref.py), every container every tick; the
    scenario provably goes through merges, prunes, the > 50 % abort and ticks without an approval."""
    from osep_path_planner_b200 import GSkel, GSkelConfig
    gr, _, sc = _pyref()
    cfg = dict(gnd_th=-100.0)
    gs = GSkel(GSkelConfig(**cfg)); ref = gr.GSkelRef(**cfg)
    stages = set()
    for k, c in enumerate(sc.structure_ticks(70, seed=0, sigma=1.3, step=0.9)):
        gs.set_input(c)
        a, b = gs.gskel_run(), ref.run(c)
        assert a == b and gs.failed_stage == ref.failed_stage, "tick %d: stage %d vs %d" % (k, gs.failed_stage, ref.failed_stage)
        _compare_gskel_with_pyref(gs, ref, "tick %d" % k)
        stages.add(ref.failed_stage)
    assert ref.events["merged"] >= 10 and ref.events["pruned"] >= 10 and {0, 1, 4} <= stages, (ref.events, stages)
    gs.close()


def test_gskel_library_edge_cases_against_pyref(mods):
    """empty input, non-finite candidates, > 50 % new abort, isolated vertex -> NaN in smooth -> dropped next tick (gskel.cpp:41,68,286,414,180)"""
    from osep_path_planner_b200 import GSkel, GSkelConfig
    gr, _, _ = _pyref()
    cfg = dict(gnd_th=-100.0)
    gs = GSkel(GSkelConfig(**cfg)); ref = gr.GSkelRef(**cfg)
    gs.set_input(np.zeros((0, 3), np.float32))
    assert gs.gskel_run() is False and ref.run(np.zeros((0, 3), np.float32)) is False and gs.failed_stage == ref.failed_stage == 1
    base = np.array([[0, 0, 0], [4, 0, 0], [8, 0, 0], [12, 0, 0], [np.nan, 0, 0]], np.float32)
    far = np.array([[0, 0, 0], [4, 0, 0], [8, 0, 0], [12, 0, 0], [60, 0, 0]], np.float32)
    saw_nan = False
    for k, c in enumerate([base] * 4 + [far] * 7):
        gs.set_input(c)
        a, b = gs.gskel_run(), ref.run(c)
        assert a == b and gs.failed_stage == ref.failed_stage, "tick %d" % k
        _compare_gskel_with_pyref(gs, ref, "edge tick %d" % k)
        saw_nan = saw_nan or bool(np.isnan(ref.positions()).any())
    assert saw_nan and not np.isnan(ref.positions()).any()
    gs.close()


def test_gskel_graph_arena_grows_instead_of_failing(mods):
    """A global skeleton that outgrows the handle's device arena must not leave the state machine ahead of the graph
    (ADVICE round 1): the arena grows and the tick completes like on a large handle."""
    osep, po, fx = mods
    from osep_path_planner_b200 import GSkel, GSkelConfig
    gsk_small, gsk_big = GSkel(GSkelConfig(gnd_th=-100.0), capacity_vertices=16), GSkel(GSkelConfig(gnd_th=-100.0), capacity_vertices=4096)
    _, _, sc = _pyref()
    for k, c in enumerate(sc.structure_ticks(30, seed=1, sigma=1.3, step=0.9)):
        a = gsk_small.gskel_run() if gsk_small.set_input(c) is not None else False
        b = gsk_big.gskel_run() if gsk_big.set_input(c) is not None else False
        assert a == b and gsk_small.failed_stage == gsk_big.failed_stage, "tick %d" % k
        assert_bit_equal(np.ascontiguousarray(gsk_small.output_gskel()), np.ascontiguousarray(gsk_big.output_gskel()), "global skeleton tick %d" % k)
    assert len(gsk_big.output_gskel()) > 16
    gsk_small.close(); gsk_big.close()


def test_voxelgrid_index_overflow_returns_the_input_cloud(mods):
    """PCL's int32 guard: when the leaf the control law asks for makes dx*dy*dz overflow, VoxelGrid
    returns its input unchanged.  A small, widely spread cloud with a large max_points drives the leaf down until that happens;
    the library must follow the oracle through it (every filtered point its own "voxel", normals not re-normalised by the
    accumulator) to whatever stage the reference then fails or succeeds at."""
    osep, po, fx = mods
    rng = np.random.default_rng(11)
    P = rng.uniform(-28, 28, (160, 3)).astype(np.float32)
    cfg = dict(max_points=1536, min_points=50)
    o, ok_o, r, ok_g = run_both(mods, P, cfg)
    assert o.flags & 4, "the oracle did not reach the overflow guard: flags %d, leaves %s" % (o.flags, o.tap("leaf_hist"))
    assert ok_g == ok_o and r.result.status == o.failed_stage, "gpu status %d vs oracle stage %d" % (r.result.status, o.failed_stage)
    assert r.result.flags & 4 and r.result.n_downsampled == len(o.tap("pts")) == len(P)
    for tap in ["filt_idx", "leaf_hist", "nvox_hist", "normals_full", "vox_keys", "vox_count", "pts", "nrms"]:
        assert_bit_equal(r.tap(tap), o.tap(tap), tap)
    if ok_o:
        assert_bit_equal(np.ascontiguousarray(r.output_vertices()), o.tap("skelver"), "vertices")
    r.close()
