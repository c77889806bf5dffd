"""BASELINE.json configs[4]: large accumulated map, spatially tiled across the GPUs of one box, DEVICE RESIDENT: halo
selection kernel + NCCL send/recv into the tile buffer, one frame per tile, owned-vertex mask on the device, NCCL
all-gather of the vertices, seam graph.  Launch with torchrun:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      tools/tiled_bench.py [total_points] [halo_m | auto]
Every rank builds the same synthetic 4-turbine map (seeded), keeps its slab and uploads it ONCE (the accumulated map
lives in HBM: SURVEY.md 8f-4); the timed region starts with the slab resident."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from osep_path_planner_b200 import fixtures as fx, tiled


def build_map(total):
    per = total // 4
    parts = []
    for k, (ox, oy) in enumerate([(0, 0), (60, 0), (0, 60), (60, 60)]):
        parts.append(fx.turbine(per, seed=100 + k) + np.array([ox, oy, 0], np.float32))
    return np.concatenate(parts).astype(np.float32)


def run(total, halo_arg, reps=3):
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    P = build_map(total)
    # partition once, when the map is laid out: recursive coordinate bisection, every cut where the fewest points lie within a
    # halo width of the cut plane (equal-count slabs along one axis are several slabs thinner than the halo on this map)
    boxes, owner = tiled.rcb_boxes(P, world, 8.0)
    own = torch.from_numpy(np.ascontiguousarray(P[owner == rank])).cuda()
    del P
    halo = None if halo_arg == "auto" else float(halo_arg)
    runner = tiled.TileRunner(local)
    res, walls = None, []
    for rep in range(reps):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        res = runner.run(own, halo=halo, box=boxes[rank])
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        walls.append(time.perf_counter() - t0)
    info = {"rank": rank, "ok": bool(res["ok"]), "status": int(res["status"]), "own": int(own.shape[0]), "tile": res["tile_points"], "halo_points": res["halo_points"],
            "halo_m": round(res["halo_m"], 3), "halo_bytes": res["halo_bytes"], "frame_gpu_ms": round(res["gpu_ms"], 3), "ms": {k: round(v, 3) for k, v in res["ms"].items()},
            "M": int(res["M"]), "owned_vertices": int(res["local"].shape[0]), "peers": res["exchange"]["peers"]}
    infos = [None] * world
    if world > 1:
        dist.all_gather_object(infos, info)
    else:
        infos = [info]
    out = None
    if rank == 0:
        ex_ms = max(i["ms"]["leaf_select_exchange"] for i in infos)
        hb = sum(i["halo_bytes"] for i in infos)
        out = {"config": "tiled map, %d points, %d GPU(s), halo %s, recursive-bisection boxes, device resident" % (total // 4 * 4, world, halo_arg),
               "seconds_per_map": walls[-1], "seconds_per_map_runs": walls, "points_per_s": (total // 4 * 4) / walls[-1],
               "tile_points": [i["tile"] for i in infos], "own_points": [i["own"] for i in infos], "halo_m": [i["halo_m"] for i in infos],
               "halo_bytes_total": hb, "exchange_ms_max": ex_ms,
               "halo_gbs_vs_nvlink": {"achieved_GBps_per_rank": (hb / max(world, 1)) / max(ex_ms, 1e-6) / 1e6, "nvlink5_GBps_per_direction": 900.0,
                                      "note": "bytes received per rank / (leaf pass + selection + count all-gather + send/recv) time: latency-bound at these sizes"},
               "tile_frame_gpu_ms": [i["frame_gpu_ms"] for i in infos], "stage_ms": [i["ms"] for i in infos], "tile_ok": [i["ok"] for i in infos],
               "tile_status": [i["status"] for i in infos], "owned_vertices": [i["owned_vertices"] for i in infos], "peers": [i["peers"] for i in infos],
               "all_vertices": int(res["vertices"].shape[0]), "seam_graph_edges": int(res["graph"]["n_edges"])}
    runner.close()
    return out


if __name__ == "__main__":
    total = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
    halo_arg = sys.argv[2] if len(sys.argv) > 2 else "auto"
    world, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    o = run(total, halo_arg)
    if o is not None:
        print(json.dumps(o))
    if world > 1:
        dist.barrier(); dist.destroy_process_group()
