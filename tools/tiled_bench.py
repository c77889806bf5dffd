"""BASELINE.json configs[4]: large accumulated map, spatially tiled across the GPUs of one box with halo
exchange (NCCL send/recv) and an all-gather of the skeleton vertices.  Launch with torchrun:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      tools/tiled_bench.py [total_points]
Every rank builds the same synthetic 4-turbine map (seeded), keeps its slab, and runs the tiled pipeline."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import osep_path_planner_b200 as osep
from osep_path_planner_b200 import fixtures as fx, tiled

total = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
per = total // 4
parts = []
for k, (ox, oy) in enumerate([(0, 0), (60, 0), (0, 60), (60, 60)]):
    parts.append(fx.turbine(per, seed=100 + k) + np.array([ox, oy, 0], np.float32))
P = np.concatenate(parts).astype(np.float32)
axis, bounds = tiled.split_axis_and_bounds(P, world)
own = np.ascontiguousarray(tiled.slab_of(P, axis, bounds, rank))
halo = float(sys.argv[2]) if len(sys.argv) > 2 else 15.0     # ~10 x the converged per-tile leaf (similarity radius, rosa.cpp:212)
r = None
info = {}

def skeletonize(tile):
    global r
    if r is None:          # sized once the tile (own slab + halos) is known
        r = osep.Rosa(osep.RosaConfig(pts_dist_lim=1.0e4), device=local, capacity_points=int(len(tile) * 1.05) + 1024)
        r.set_profile(True)
    r.set_input(tile)
    ok = r.rosa_run()
    info.update(ok=ok, status=r.result.status, M=r.result.n_downsampled, gpu_ms=r.result.gpu_ms, n=len(tile), leaf=r.result.leaf_size, stage_ms={k: round(v, 3) for k, v in r.stage_ms().items()}, knn_stats=[int(x) for x in r.tap("knn_stats")])
    return np.ascontiguousarray(r.output_vertices())

def sync():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()

for rep in range(3):
    sync(); t0 = time.perf_counter()
    res = tiled.tiled_skeleton(own, (bounds[rank], bounds[rank + 1]), axis, halo, skeletonize)
    sync(); dt = time.perf_counter() - t0
allv = np.concatenate([v for v in res["vertices"] if len(v)]) if sum(len(v) for v in res["vertices"]) else np.zeros((0, 3), np.float32)
g = osep.build_graph(np.ascontiguousarray(allv), 3.0, device=local) if len(allv) else {"n_edges": 0}
infos = [None] * world
if world > 1:
    dist.all_gather_object(infos, info)
else:
    infos = [info]
if rank == 0:
    print(json.dumps({"config": "tiled map, %d points, %d GPU(s), halo %.2f m along axis %d" % (len(P), world, halo, axis), "seconds_per_map": dt,
                      "points_per_s": len(P) / dt, "tile_points": [i["n"] for i in infos], "tile_gpu_ms": [i["gpu_ms"] for i in infos],
                      "tile_ok": [i["ok"] for i in infos], "stage_ms_rank0": infos[0]["stage_ms"], "stage_ms_all": [i["stage_ms"] for i in infos], "knn_stats_all": [i.get("knn_stats") for i in infos], "owned_vertices": [int(c) for c in res["counts"]], "seam_graph_edges": int(g["n_edges"])}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
