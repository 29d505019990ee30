"""Device timing helpers (CUDA events, max over ranks) -- SURVEY.md D10, BASELINE.md section 5."""
from __future__ import annotations

import statistics
import time
from typing import Callable, List

import torch
import torch.distributed as dist

_L2_FLUSH = {}


def l2_flush(device=None, nbytes: int = 256 << 20) -> None:
    """Write a buffer larger than the 126 MB L2 so the next kernel starts cold."""
    device = torch.device(device or "cuda")
    key = (device.index, nbytes)
    buf = _L2_FLUSH.get(key)
    if buf is None:
        buf = torch.empty(nbytes // 4, dtype=torch.int32, device=device)
        _L2_FLUSH[key] = buf
    buf.zero_()


def time_cuda(
    fn: Callable[[], object],
    steps: int = 20,
    warmup: int = 3,
    flush_l2: bool = False,
    group=None,
    barrier: bool = True,
) -> dict:
    """Time ``fn`` with CUDA events on the current stream.

    Returns per-call milliseconds: list, median, mean, and (if a process group is up) the MAX
    over ranks of the total.
    """
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    if multi and barrier:
        dist.barrier(group)
    torch.cuda.synchronize()
    t_total0 = torch.cuda.Event(enable_timing=True)
    t_total1 = torch.cuda.Event(enable_timing=True)
    t_total0.record()
    for i in range(steps):
        if flush_l2:
            l2_flush()
        starts[i].record()
        fn()
        ends[i].record()
    t_total1.record()
    torch.cuda.synchronize()
    per = [s.elapsed_time(e) for s, e in zip(starts, ends)]
    total = sum(per)
    if multi:
        t = torch.tensor([total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        total_max = float(t.item())
        tm = torch.tensor([statistics.median(per)], device="cuda", dtype=torch.float64)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX, group=group)
        med_max = float(tm.item())
    else:
        total_max = total
        med_max = statistics.median(per)
    return {
        "per_call_ms": per,
        "median_ms": statistics.median(per),
        "median_ms_max_over_ranks": med_max,
        "mean_ms": total / steps,
        "total_ms_max_over_ranks": total_max,
        "ms_per_step": total_max / steps,
        "wall_loop_ms": t_total0.elapsed_time(t_total1),
    }


def time_host(fn: Callable[[], object], steps: int = 5, warmup: int = 1) -> dict:
    """Host-clock timing for CPU paths."""
    for _ in range(warmup):
        fn()
    per: List[float] = []
    for _ in range(steps):
        t0 = time.perf_counter()
        fn()
        per.append((time.perf_counter() - t0) * 1e3)
    return {"per_call_ms": per, "median_ms": statistics.median(per), "ms_per_step": sum(per) / steps}
