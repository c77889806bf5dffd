"""Multi-GPU plumbing for the batched-frames configuration (BASELINE.json configs[3]).

The path shards by independent frames: frame b belongs to rank b // ceil(B / P) (contiguous
blocks).  There is NO data-path collective: each rank runs whole frames on its
own GPU.  The only exchange is the final gather of the tiny per-frame results (vertex counts
+ capacity-padded vertex buffers) with torch.distributed all_gather (NCCL on GPUs, gloo in
the CPU tests) and the max-over-ranks reduction of the timing.
"""
import numpy as np
import torch
import torch.distributed as dist


def frames_of_rank(n_frames, rank, world):
    per = -(-n_frames // world)
    return list(range(min(rank * per, n_frames), min((rank + 1) * per, n_frames)))


def _dev():
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def gather_frame_results(local, n_frames, cap=None):
    """local: {frame id: (V_b, 3) float32}.  Returns (counts[n_frames], [vertices_b]) on every rank.
    The buffers are sized from the largest vertex count over all ranks (one all_reduce(MAX) of a scalar), so no frame is
    ever truncated; `cap` only sets a minimum row count (fixed-shape buffers across calls)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    per = -(-n_frames // world)
    rank = dist.get_rank() if dist.is_initialized() else 0
    vmax = max([len(v) for v in local.values()] + [1])
    if world > 1:
        t = torch.tensor([vmax], dtype=torch.int64, device=_dev())
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        vmax = int(t.item())
    rows = max(vmax, cap or 0)
    cnt = torch.zeros(per, dtype=torch.int32)
    buf = torch.zeros(per, rows, 3, dtype=torch.float32)
    for b, v in local.items():
        k = b - rank * per
        n = len(v)
        cnt[k] = n
        if n:
            buf[k, :n] = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32))
    if world == 1:
        counts = cnt.numpy()[:n_frames]
        return counts, [buf[b, :counts[b]].numpy() for b in range(n_frames)]
    d = _dev()
    cnt_all = [torch.zeros_like(cnt, device=d) for _ in range(world)]
    buf_all = [torch.zeros_like(buf, device=d) for _ in range(world)]
    dist.all_gather(cnt_all, cnt.to(d))
    dist.all_gather(buf_all, buf.to(d))
    counts = torch.cat(cnt_all).cpu().numpy()[:n_frames]
    allbuf = torch.cat(buf_all).cpu().numpy()
    return counts, [allbuf[b, :counts[b]] for b in range(n_frames)]


def max_over_ranks(x):
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=_dev())
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
