#!/usr/bin/env python
"""bench.py -- skeleton frames/s and ms/frame at 100k points (BASELINE.json metric).

A step = one pass of the hot path (Rosa::rosa_run: point cloud -> skeleton vertices, plus the
skeleton graph) over one synthetic 100k-point turbine frame (BASELINE.json configs[1]).
  value : frames/s with the frame already resident in HBM (device entry point of the C ABI),
          CUDA events on the library's stream, L2 flushed between steps.
  e2e   : the same frames through the reference-facing call (osep_rosa_run: HOST pinned buffer
          -> H2D -> 7 stages -> D2H of the vertices), wall clock around the synchronous call.
N > 1 (torchrun): one process per GPU, each rank runs its own frames (independent frames are
the unit the path shards by; no data-path collective), value = frames of all ranks / max time.
Extra blocks on the same JSON line: `batched` (configs[3]: 256 distinct frames from pinned host
buffers through osep_rosa_run_batch, results gathered over NCCL inside the timed region),
`stream` (configs[2] in miniature: frames with a different n each, p50/p99 latency) and, for
N >= 2, `tiled` (configs[4]: one map split into spatial tiles with a halo exchange).
--impl reference: the CPU restatement of the reference (oracle, faithful mode, the reference
is single threaded) on the host cores, same metric.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np   # noqa: E402

METRIC = "skeleton_frames_per_sec_100k_pts"
WORKLOAD = "single frame, 100k-point synthetic wind-turbine cloud (tower + 3 blades, sigma 0.02 m), default RosaConfig"


def env_int(k, d):
    try:
        return int(os.environ.get(k, d))
    except ValueError:
        return d


def algorithmic_bytes(n_in, n_filt, L, M, V):
    """16 N + N' (164 + 12 (L-1)) + 48 M + 12 V bytes per frame."""
    return 16 * n_in + n_filt * (164 + 12 * (L - 1)) + 48 * M + 12 * V


class ClockSampler:
    Q = "timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows = []
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([time.time()] + [x.strip() for x in line.split(",")][1:])

    def mark(self):
        """start of the timed region: only samples taken after this call are reported"""
        self.t_mark = time.time()

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(2)
        except Exception:
            self.p.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t_mark = getattr(self, "t_mark", 0.0)
        rows = [r[1:] for r in self.rows if r[0] >= t_mark] or [r[1:] for r in self.rows]
        for r in rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except ValueError:
                continue
            for nme, val in zip(names, r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def time_oracle(P, frames, faithful=True):
    from oracle import pyoracle as po
    kw = dict(faithful_normals=1, faithful_order=1, libm_cbrt=1) if faithful else {}
    o = po.RosaOracle(**kw)
    ts, stages = [], None
    for _ in range(frames):
        t0 = time.perf_counter()
        ok = o.run(P)
        v = o.tap("skelver")
        if ok:
            po.GraphOracle(v, 3.0)
        ts.append(time.perf_counter() - t0)
        stages = o.stage_ms()
    return ts, stages, int(len(o.tap("skelver")))


def time_oracle_all_cores(P, rounds=1):
    """Process-parallel CPU throughput: one independent frame per host core, all cores busy.
    The oracle is a ctypes call (GIL released), so threads are enough."""
    from concurrent.futures import ThreadPoolExecutor
    nproc = os.cpu_count() or 1
    t0 = time.perf_counter()
    with ThreadPoolExecutor(nproc) as ex:
        list(ex.map(lambda _: time_oracle(P, rounds), range(nproc)))
    dt = time.perf_counter() - t0
    return {"value": nproc * rounds / dt, "unit": "frames/s", "cores": nproc,
            "sample": "%d independent 100k-point frames, one per core (the batched workload of configs[3]); a single frame cannot use more than one core" % (nproc * rounds)}


def run_reference(args, rank, world):
    """Reference arm: the reference's own CPU algorithm (restated; the C++ reference cannot be
    compiled here -- no PCL/Eigen/FLANN/ROS 2) on the host cores.  Single thread: the reference
    node is single threaded end to end (rosa.hpp:80, rclcpp::spin single-threaded executor)."""
    if rank != 0:
        return
    from osep_path_planner_b200 import fixtures as fx
    P = fx.turbine(args.points, seed=0)
    time_oracle(P, max(1, min(args.warmup, 2)))
    ts, stages, V = time_oracle(P, args.steps)
    ms = 1e3 * float(np.mean(ts))
    fps = 1e3 / ms
    line = {"metric": METRIC, "value": fps, "unit": "frames/s", "impl": "reference", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "points": args.points, "n_vertices": V},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": 1, "kind": "port",
                             "sample": "%d full frames, oracle faithful mode (reference summation orders, PCL trig eigen33, libm cbrt), 1 thread" % args.steps,
                             "stage_ms": dict(zip(["preprocess", "rosa_init", "similarity_neighbor_extraction", "drosa", "dcrosa", "vertex_sampling", "vertex_smooth"], stages)),
                             "nproc": os.cpu_count(), "all_cores": time_oracle_all_cores(P)},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import osep_path_planner_b200 as osep
    from osep_path_planner_b200 import fixtures as fx, shard

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # keep stdout = the one JSON line: NCCL prints its version banner to fd 1 when the communicator comes up
        sys.stdout.flush()
        saved = os.dup(1); os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier(); torch.cuda.synchronize()
        finally:
            sys.stdout.flush(); os.dup2(saved, 1); os.close(saved)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # weak scaling over independent frames: every rank runs its own replica of the named workload (configs[1], seed 0)
    P4 = fx.to_xyz4(fx.turbine(args.points, seed=0))
    n = P4.shape[0]
    host = torch.from_numpy(P4).pin_memory()
    dev = host.cuda(non_blocking=False)
    r = osep.Rosa(device=local_rank, capacity_points=n)
    r.set_graph(True, 3.0)             # frame -> skeleton graph in one launch sequence: graph_adj + mst run inside the frame
    stream = torch.cuda.ExternalStream(r.stream, device=local_rank)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")      # > 126 MB L2

    def step_device():
        r.run_device(dev.data_ptr(), n, 4)

    clocks = ClockSampler(local_rank)      # started early (nvidia-smi takes a while to come up); samples before mark() are dropped
    # ---- pre-warm: a fresh box needs a few hundred ms of work before clocks / caches / the CUDA graph settle; the W
    # untimed warm-up steps of the contract follow
    t_pw = time.perf_counter()
    while time.perf_counter() - t_pw < args.prewarm:
        step_device(); assert r.finish()
    # ---- warm-up
    for _ in range(args.warmup):
        step_device(); assert r.finish(), "frame failed: status %d" % r.result.status
        g = r.graph(3.0)
    res0 = r.result
    launches0 = r.launches

    # ---- timed: device-resident input, CUDA events on the library's stream, L2 flushed between steps
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    stage_acc = None
    clocks.mark()
    barrier()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        with torch.cuda.stream(stream):
            flush.zero_()                       # L2 flush, outside the per-step events
            ev[k][0].record(stream)
        step_device()                           # 7 Rosa stages + graph_adj + mst + result copies: one launch sequence (CUDA graph)
        with torch.cuda.stream(stream):
            ev[k][1].record(stream)
        ok = r.finish()
        assert ok
        g = r.graph(3.0)                        # host views (CSR, joints, leafs) of the graph the frame built on the device
    timeline_us = r.timeline()                  # kernel entry times of the last timed frame (normal overlapped run, L2 cold)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = r.launches - launches0
    # per-stage device times: a few extra frames with the library's stage events on (this serialises the two
    # streams a frame normally overlaps, so it is kept out of the timed region)
    r.set_profile(True)
    n_prof = 5
    for _ in range(n_prof):
        step_device(); assert r.finish()
        sm = r.tap("stage_ms")
        stage_acc = sm if stage_acc is None else stage_acc + sm
    r.set_profile(False)
    dev_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = shard.max_over_ranks(sum(dev_ms))
    ms_per_step = total_ms / args.steps
    value = world * args.steps / (total_ms * 1e-3)

    # ---- e2e: host pinned buffer -> osep_rosa_run (H2D + frame + D2H vertices), synchronous call
    import ctypes as C
    L = osep.load()
    res = osep.RosaResult()
    for _ in range(max(1, args.warmup // 2)):
        L.osep_rosa_run(r._h, C.c_void_p(host.data_ptr()), n, 4, C.byref(res))
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        rc = L.osep_rosa_run(r._h, C.c_void_p(host.data_ptr()), n, 4, C.byref(res))
        assert rc == 0
        g = r.graph(3.0)
    barrier()
    e2e_s = shard.max_over_ranks(time.perf_counter() - t0)
    clk = clocks.stop()
    e2e_value = world * args.steps / e2e_s
    d2h = 12 * 1024 + 256 + int(g["n_vertices"]) * 40 + int(g["n_edges"]) * 12 + 16     # vertices staging (first 768 float4) + frame state + graph lists

    stage_ms = dict(zip(r.STAGE_NAMES, (stage_acc / n_prof).tolist()))
    stage_ms.pop("-", None)
    # dominant kernel group, timed live with the library's CUDA events on its stream
    dom = max(stage_ms, key=stage_ms.get)
    alg = algorithmic_bytes(n, res0.n_filtered, res0.n_passes, res0.n_downsampled, res0.n_vertices)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)"
    frame_gbs = alg / (ms_per_step * 1e-3) / 1e9
    M_, V_, S_ = res0.n_downsampled, res0.n_vertices, res0.n_simi
    Wn = (M_ + 31) // 32
    # dominant kernel = k_drosa_fused (one launch per frame): algorithmic bytes = its operands read once + per-iteration state
    # written/read once: pts + nraw + vnd (48 M), similarity bit matrix (4 M Wn), surf lists (4 nbk M), per-iteration
    # orientation/weight exchange (2 x 16 M x (niter+1)), outputs (pset 16 M + flags).  Everything is L2 resident.
    nit = 6
    drosa_bytes = 48 * M_ + 4 * M_ * Wn + 4 * 10 * M_ + 2 * 16 * M_ * (nit + 1) + 16 * M_
    drosa_ms = stage_ms["drosa_orient_smooth"]
    drosa_gbs = drosa_bytes / (drosa_ms * 1e-3) / 1e9
    knn_ms = stage_ms["knn_normals"]
    knn_bytes = 28 * res0.n_filtered                       # SURVEY 8(d) phase (3): read 16 N', write 12 N'
    roofline = {"bound": "hbm", "kernel": "k_drosa_fused (persistent cooperative kernel: 6 orientation iterations + position step, %.0f %% of the frame)" % (100 * drosa_ms / ms_per_step),
                "achieved": drosa_gbs, "peak": peak, "unit": "GB/s", "frac": drosa_gbs / peak,
                "traffic": 519168, "traffic_source": "profiles/r2_s7_ncu_full_kernels.md (dram__bytes_read.sum + dram__bytes_write.sum of k_drosa_fused, one ncu --set full capture, cold cache)",
                "peak_source": peak_src, "algorithmic_bytes_per_launch": drosa_bytes, "ms_per_launch": drosa_ms,
                "note": "latency-bound by construction: ~236 dependent 3x3 eigen solves per Gauss-Seidel chain (rosa.cpp:295-308), six chains as overlapping wavefronts, "
                        "one warp per chain at ~4 cycles per dependent instruction (ncu: 7.8 % issue-active, 90 % L2 hit); the HBM fraction is reported for the record, the signal is ms_per_launch",
                "frame": {"algorithmic_bytes": alg, "achieved": frame_gbs, "frac": frame_gbs / peak,
                          "note": "SURVEY.md 8(d) formula; 26 MB of compulsory traffic = 4 us at peak, a single frame is latency-bound (SURVEY.md H3)"},
                "knn_normals_stage": {"ms": knn_ms, "algorithmic_bytes": knn_bytes, "achieved": knn_bytes / (knn_ms * 1e-3) / 1e9 if knn_ms > 0 else None,
                                      "frac": (knn_bytes / (knn_ms * 1e-3) / 1e9) / peak if knn_ms > 0 else None,
                                      "note": "thread-per-query scan + fall-backs (ncu: 16.6 M warp instructions = 168 per query in the scan, 43 % issue-active at 20 warps per SM), L1/L2 resident"}}

    # ---- BASELINE.json configs[3]: 256 DISTINCT independent frames (seeds 0..255, random yaw / stand-off), 256 / N per rank,
    # through the host entry point osep_rosa_run_batch from pinned host buffers (H2D inside the timed region), per-frame
    # vertices gathered over NCCL inside the timed region.  All runs are listed; the worst one is reported next to the median.
    batched = None
    if args.batch > 0:
        nb = max(1, args.batch // world)
        r.set_lanes(args.lanes)
        frames_h = [torch.from_numpy(fx.to_xyz4(fx.batch_frame(rank * nb + b, n=args.points))).pin_memory() for b in range(nb)]
        ptrs = [f.data_ptr() for f in frames_h]
        cnts = [int(f.shape[0]) for f in frames_h]

        def run_batch_host():
            B = len(ptrs)
            pa = (C.c_void_p * B)(*ptrs); na = (C.c_int * B)(*cnts)
            resv = (osep.RosaResult * B)()
            rc = L.osep_rosa_run_batch(r._h, pa, na, B, 4, resv)
            assert rc == 0 and all(x.status == 0 for x in resv), "batched run failed: %d" % rc
            local = {}
            for b in range(B):
                pp = C.POINTER(C.c_float)(); nv = C.c_int()
                L.osep_rosa_batch_vertices(r._h, b, C.byref(pp), C.byref(nv))
                local[rank * nb + b] = np.ctypeslib.as_array(pp, shape=(3, nv.value)).T if nv.value else np.zeros((0, 3), np.float32)
            counts, _ = shard.gather_frame_results(local, nb * world)
            return resv, counts
        run_batch_host()                                                         # warm-up: every lane captures its CUDA graph
        runs = []
        for _ in range(args.batch_runs):
            barrier()
            t0 = time.perf_counter()
            bres, bcounts = run_batch_host()
            barrier()
            runs.append(shard.max_over_ranks(time.perf_counter() - t0))
        bt = float(np.median(runs)); total = nb * world
        b_alg = float(np.mean([algorithmic_bytes(cnts[b], bres[b].n_filtered, bres[b].n_passes, bres[b].n_downsampled, bres[b].n_vertices) for b in range(nb)]))
        b_gbs = b_alg * total / bt / 1e9
        batched = {"frames": total, "distinct_frames": total, "frames_per_rank": nb, "lanes_per_gpu": args.lanes,
                   "frames_per_s": total / bt, "frames_per_s_worst_run": total / max(runs), "frames_per_s_runs": [total / t for t in runs],
                   "max_over_median_run_time": max(runs) / bt, "ms_per_frame_amortised": 1e3 * bt / nb,
                   "e2e": {"value": total / bt, "unit": "frames/s", "h2d_bytes_per_frame": int(np.mean(cnts)) * 16, "d2h_bytes_per_frame": 12 * 1024 + 256,
                           "gather": "per-frame vertices all_gathered over NCCL inside the timed region" if world > 1 else "single rank: no gather"},
                   "roofline": {"bound": "hbm", "algorithmic_bytes_per_frame": b_alg, "achieved": b_gbs, "peak": peak * world, "unit": "GB/s", "frac": b_gbs / (peak * world),
                                "note": "SURVEY.md 8(d) bytes per frame x frames/s over all ranks / (measured HBM peak x ranks): the configuration where the HBM fraction is meaningful"},
                   "mean_vertices_per_frame": float(np.mean(bcounts)),
                   "note": "pinned HOST frames through osep_rosa_run_batch, %d frames in flight per GPU, wall clock max over ranks, %d runs" % (args.lanes, args.batch_runs)}
        r.set_lanes(1)
        del frames_h
    # ---- BASELINE.json configs[2] in miniature: a stream of C3 frames whose point count changes every frame (a real LiDAR never
    # repeats n), Rosa (host entry point) + TF + GSkel tick per frame; per-frame latency percentiles.  Rank 0 only.
    stream = None
    if args.stream > 0 and rank == 0:
        from osep_path_planner_b200 import GSkel, GSkelConfig
        rng = np.random.default_rng(7)
        sf = []
        for f in range(args.stream):
            P, Rm, tv = fx.stream_frame(f * 5, n=50000)
            keep_n = int(rng.integers(42000, 50001))
            sf.append((torch.from_numpy(fx.to_xyz4(P[:keep_n])).pin_memory(), Rm, tv))
        gsk = GSkel(GSkelConfig(gnd_th=-100.0), device=local_rank)
        lat_r, lat_g, replays0 = [], [], int(r.tap("graph_replays")[0])
        for wu in range(3):
            L.osep_rosa_run(r._h, C.c_void_p(sf[wu][0].data_ptr()), int(sf[wu][0].shape[0]), 4, C.byref(res))
        for buf, Rm, tv in sf:
            t0 = time.perf_counter()
            rc = L.osep_rosa_run(r._h, C.c_void_p(buf.data_ptr()), int(buf.shape[0]), 4, C.byref(res))
            t1 = time.perf_counter()
            if rc == 0:
                v = np.ascontiguousarray(r.output_vertices_now())
                gsk.set_input((v.astype(np.float64) @ Rm.T + tv).astype(np.float32)); gsk.gskel_run()
            lat_r.append(1e3 * (t1 - t0)); lat_g.append(1e3 * (time.perf_counter() - t1))
        stream = {"frames": len(sf), "points_per_frame": [int(min(x[0].shape[0] for x in sf)), int(max(x[0].shape[0] for x in sf))],
                  "distinct_point_counts": len(set(int(x[0].shape[0]) for x in sf)),
                  "rosa_ms_p50": float(np.percentile(lat_r, 50)), "rosa_ms_p99": float(np.percentile(lat_r, 99)), "rosa_ms_max": float(np.max(lat_r)),
                  "gskel_ms_p50": float(np.percentile(lat_g, 50)), "gskel_ms_p99": float(np.percentile(lat_g, 99)),
                  "graph_replays": int(r.tap("graph_replays")[0]) - replays0, "global_vertices": int(len(gsk.output_gskel())),
                  "note": "host pinned frame -> osep_rosa_run (H2D + frame + D2H) wall clock per frame; every frame has a different n and replays the same CUDA graph"}
        gsk.close()
    # ---- BASELINE.json configs[4]: one accumulated map, spatially tiled across the N GPUs, device resident (2.5 M points per GPU:
    # 20 M at N = 8).  Halo selection kernel + NCCL send/recv + one frame per tile + NCCL all-gather of the owned vertices + seam graph.
    tiled_blk = None
    if world > 1 and args.tiled_points > 0:
        import importlib.util
        spec = importlib.util.spec_from_file_location("tiled_bench", os.path.join(ROOT, "tools", "tiled_bench.py"))
        tb = importlib.util.module_from_spec(spec); spec.loader.exec_module(tb)
        tiled_blk = tb.run(args.tiled_points * world, "auto", reps=3)
    if world > 1:
        counts, _ = shard.gather_frame_results({rank: np.ascontiguousarray(r.output_vertices())}, world, cap=1024)
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        ts, cstages, _ = time_oracle(fx.turbine(args.points, seed=0), args.cpu_frames)
        cpu_ms = 1e3 * float(np.mean(ts))
        cpu_baseline = {"value": 1e3 / cpu_ms, "unit": "frames/s", "cores": 1, "kind": "port",
                        "sample": "%d full 100k-point frames, oracle faithful mode, 1 thread (the reference is single-threaded)" % args.cpu_frames,
                        "ms_per_frame": cpu_ms, "nproc": os.cpu_count(), "all_cores": time_oracle_all_cores(fx.turbine(args.points, seed=0))}
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD, "points": n, "n_filtered": res0.n_filtered, "voxel_passes": res0.n_passes, "n_downsampled": res0.n_downsampled,
                           "n_vertices": res0.n_vertices, "graph_edges": int(g["n_edges"]), "l2": "flushed between steps (256 MiB memset, outside the per-step events)", "prewarm_s": args.prewarm,
                           "parallelism": "1 frame per GPU per step; ranks = independent frames (no collective on the data path)"},
                "roofline": roofline, "cpu_baseline": cpu_baseline,
                "e2e": {"value": e2e_value, "unit": "frames/s", "ms_per_step": 1e3 * e2e_s / args.steps, "h2d_bytes_per_step": int(n * 16), "d2h_bytes_per_step": d2h},
                "gpu_launches": int(launches), "stage_ms": stage_ms, "timeline_us": timeline_us, "clocks": clk, "wall_ms_per_step_incl_flush": 1e3 * t_wall / args.steps, "batched": batched, "stream": stream, "tiled": tiled_blk}
        print(json.dumps(line), flush=True)
    r.close()
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--points", type=int, default=100000)
    ap.add_argument("--cpu-frames", type=int, default=20, help="frames of the bounded CPU-baseline sample (~0.4 s each)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--batch", type=int, default=256, help="total frames of the batched measurement, configs[3] (split over the ranks; 0 = skip)")
    ap.add_argument("--batch-runs", type=int, default=5, help="timed repetitions of the batched measurement (all are listed)")
    ap.add_argument("--stream", type=int, default=200, help="frames of the varying-n stream measurement, configs[2] in miniature (0 = skip)")
    ap.add_argument("--tiled-points", type=int, default=2500000, help="points PER GPU of the tiled-map measurement, configs[4] (N >= 2 only; 0 = skip)")
    ap.add_argument("--prewarm", type=float, default=0.5, help="seconds of untimed frames before the W warm-up steps")
    ap.add_argument("--lanes", type=int, default=16, help="concurrent frames (streams) per GPU in the batched measurement (12 drosa kernels of 12 CTAs are co-resident; the extra lanes keep N-scale work in flight)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    import __graft_entry__ as ge
    if rank == 0 and not os.path.exists(os.path.join(ROOT, "osep_path_planner_b200", "libosep_skel.so")):
        ge.build()
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
