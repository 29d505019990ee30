"""Sequence-length sweep (BASELINE.json config #5): 4K-1M global tokens at the launched world size, fused
in-kernel tree combine vs the reference-structured NCCL path.  Run under torchrun; rank 0 appends one JSON
line per (mode, seq, impl) to --out (resumable: finished keys are skipped).

Latency: CUDA events, median of --steps calls after warm-up, MAX over ranks.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "baseline"))

import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/sweep.jsonl")
    ap.add_argument("--seqs", type=int, nargs="*", default=[4096, 16384, 65536, 131072, 262144, 1048576])
    ap.add_argument("--heads", type=int, default=32)
    ap.add_argument("--kv-heads", type=int, default=None)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--prefill-sq", type=int, default=4096, help="query-block length of the chunked-prefill rows")
    ap.add_argument("--modes", nargs="*", default=["decode", "prefill"])
    a = ap.parse_args()

    import tree_attention_b200 as ta
    import nccl_minfix
    from tree_attention_b200.utils.timing import time_cuda

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    ta.setup(rank, world, local_rank=int(os.environ.get("LOCAL_RANK", str(rank))))
    dev = torch.device("cuda", torch.cuda.current_device())
    hkv = a.kv_heads or a.heads
    done = set()
    if rank == 0 and os.path.exists(a.out):
        for line in open(a.out):
            try:
                d = json.loads(line)
                done.add((d["mode"], d["seq_global"], d["impl"]))
            except Exception:
                pass
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass

    def emit(rec):
        if rank == 0:
            os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
            with open(a.out, "a") as f:
                f.write(json.dumps(rec) + "\n")
            print(json.dumps(rec), flush=True)

    for mode in a.modes:
        for S in a.seqs:
            if S % world:
                continue
            s_local = S // world
            sq = 1 if mode == "decode" else min(a.prefill_sq, S)
            if mode == "prefill" and s_local * hkv * a.dim * 4 > 40e9:
                continue
            q, k, v = ta.make_data((1, a.heads, s_local, a.dim), rank, dev, dtype=torch.bfloat16, sq=sq, num_kv_heads=hkv, log=False)
            scale = a.dim ** -0.5
            impls = {
                "fused": lambda: ta.tree_attention(q, k, v, softmax_scale=scale, backend="fused"),
            }
            if world > 1:
                impls["symm_combine"] = lambda: ta.tree_attention(q, k, v, softmax_scale=scale, backend="symm")
                impls["own_kernel+nccl_allreduce3"] = lambda: ta.tree_attention(q, k, v, softmax_scale=scale, backend="nccl", schedule="allreduce3")
                if mode == "decode" or sq * s_local * a.heads * 2 < 8e9:
                    impls["nccl_minfix_reference_structure"] = lambda: nccl_minfix.tree_decode_minfix(q, k, v, scale)
            for name, fn in impls.items():
                key = (mode, S, name)
                skip = torch.tensor([1 if key in done else 0], device=dev)
                if world > 1:
                    dist.broadcast(skip, 0)
                if int(skip.item()):
                    continue
                try:
                    t = time_cuda(fn, steps=a.steps, warmup=a.warmup)
                    ms = t["median_ms_max_over_ranks"]
                    rec = {"mode": mode, "seq_global": S, "world": world, "impl": name, "sq": sq, "heads": a.heads,
                           "kv_heads": hkv, "dim": a.dim, "latency_us": ms * 1e3, "tokens_per_s": S / (ms * 1e-3) if mode == "decode" else sq / (ms * 1e-3)}
                    if mode == "decode":
                        kvb = 2 * hkv * s_local * a.dim * 2
                        rec["hbm_gbs_per_gpu"] = kvb / (ms * 1e-3) / 1e9
                        rec["hbm_frac_of_measured"] = rec["hbm_gbs_per_gpu"] / peaks.get("hbm_gbs", 6650.0)
                        rec["combine_payload_bytes_per_rank"] = (world - 1) * a.heads * sq * (a.dim + 4) * 4
                    else:
                        fl = 4.0 * sq * s_local * a.dim * a.heads
                        rec["tflops_per_gpu"] = fl / (ms * 1e-3) / 1e12
                        rec["frac_of_measured_cublas"] = rec["tflops_per_gpu"] / peaks.get("bf16_tflops", 1590.0)
                        rec["nvlink_bytes_pushed_per_rank"] = (world - 1) * a.heads * sq * (a.dim * 2 + 4)
                    emit(rec)
                except Exception as e:
                    emit({"mode": mode, "seq_global": S, "world": world, "impl": name, "error": f"{type(e).__name__}: {e}"[:200]})
            del q, k, v
            torch.cuda.empty_cache()
    ta.cleanup()


if __name__ == "__main__":
    main()
