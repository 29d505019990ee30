"""fp32 sum over ranks: the symmetric-memory kernel (csrc/reduce.cu, the backward's dQ reduce) vs dist.all_reduce (NCCL).
torchrun --nproc-per-node W bench_tools/bench_allreduce.py [--mib 16 256 1024]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import tree_attention_b200 as ta
from tree_attention_b200.parallel.tree import allreduce_sum

ap = argparse.ArgumentParser(); ap.add_argument("--mib", type=int, nargs="*", default=[16, 256, 1024]); a = ap.parse_args()
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
ta.setup(rank, world, local_rank=int(os.environ.get("LOCAL_RANK", rank)))
dev = torch.device("cuda", torch.cuda.current_device())
def timed(fn, steps=8, warm=3):
    for _ in range(warm): fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)
for mib in a.mib:
    n = mib << 18
    x = torch.randn(n, device=dev)
    y = allreduce_sum(x); z = x.clone(); dist.all_reduce(z)
    err = float((y - z).abs().max())
    t_own = timed(lambda: allreduce_sum(x)); zz = x.clone(); t_nccl = timed(lambda: dist.all_reduce(zz))
    bus = lambda ms: 2 * (world - 1) / world * n * 4 / (ms * 1e-3) / 1e9
    if rank == 0:
        print(json.dumps({"world": world, "MiB": mib, "own_symm_ms": t_own, "nccl_ms": t_nccl, "own_busbw_gbs": bus(t_own),
                          "nccl_busbw_gbs": bus(t_nccl), "own_over_nccl": t_nccl / t_own, "max_abs_diff": err}), flush=True)
ta.cleanup()
