"""The BASELINE.json configs beyond the headline decode step, measured through the public API:

  (a) ``fwd_128k_bf16``      full-Sq forward, seq = 128K, 32 heads, d = 128, bf16: Q replicated, KV sharded over the ranks,
                             ONE fused launch per rank and query chunk (tcgen05 flash forward + in-kernel cross-GPU combine);
  (b) ``decode_256k_mxfp8``  seq = 256K, 32 heads, block-scaled fp8 KV (e4m3 + UE8M0): decode step with both GEMMs on
                             ``tcgen05.mma.kind::mxf8f6f4.block_scale`` and the tree combine fused in;
  (c) ``gqa_1m_fwd_bwd``     seq = 1M, GQA 32q / 8kv, forward + backward (dK, dV local; dQ summed over ranks by the
                             symmetric-memory reduce), plus that reduce timed against ``dist.all_reduce``.

Each block carries latency (CUDA events, max over ranks), the BASELINE.md section 5 metrics, the fraction of the measured
roofline denominator (``MEASURED_PEAKS.json``) and a clock record sampled with nvidia-smi during the measurement.
Called from ``bench.py --heavy on`` (default on 8 GPUs) and runnable on its own under torchrun:

    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 bench_tools/configs.py --out gpurun_out/configs_8.json

Reference: generalises the one decode step of /root/reference/model.py:129-155 (seq 64 000 per rank, 16 heads) to the
configs BASELINE.json names; the reference itself has no forward for Sq > 1, no backward and no fp8.
"""
from __future__ import annotations

import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _peaks(root):
    try:
        return json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def _timed(torch, dist, fn, steps, warmup, world, dev, barrier):
    for _ in range(warmup):
        fn()
    barrier()
    t0w = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t1w = time.time()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()), (t0w, t1w)


def run_all(ta, world, rank, dev, barrier, root, only=None, sampler=None, scale_down=1):
    import torch
    import torch.distributed as dist

    from bench import ClockSampler

    peaks = _peaks(root)
    hbm = peaks.get("hbm_gbs", 6650.0)
    tf_sus = peaks.get("bf16_tflops_sustained", 1400.0)
    own_sampler = sampler is None and rank == 0
    if own_sampler:
        sampler = ClockSampler(dev.index if dev.index is not None else 0)
    out = {}

    def clocks(window):
        if sampler is None:
            return None
        c = sampler.summary(*window)
        c["source"] = "nvidia-smi during the timed region"
        return c

    def want(name):
        return only is None or name in only

    # ------------------------------------------------------------------------------------------ (a) full-Sq forward
    if want("fwd_128k_bf16"):
        try:
            S, H, D = 131072 // scale_down, 32, 128
            s_local = S // world
            g = torch.Generator(device=dev).manual_seed(77)
            q = torch.randn(1, H, S, D, device=dev, generator=g).to(torch.bfloat16)   # same seed on every rank: replicated Q
            _, k, v = ta.make_data((1, H, s_local, D), rank, dev, dtype=torch.bfloat16, seed=5, log=False)
            blk = {"seq": S, "heads": H, "head_dim": D, "dtype": "bf16", "q": "replicated", "kv_tokens_per_rank": s_local,
                   "backend": "fused: tcgen05 flash forward + in-kernel cross-GPU combine, one launch per rank per query chunk"}
            for causal in (False, True):
                fn = lambda: ta.tree_attention(q, k, v, causal=causal, backend="fused" if world > 1 else "auto")
                ms, win = _timed(torch, dist, fn, 3, 1, world, dev, barrier)
                flops = 4.0 * S * s_local * D * H * (0.5 if causal else 1.0)
                tf = flops / (ms * 1e-3) / 1e12
                key = "causal" if causal else "full"
                blk[key] = {
                    "ms": ms, "tokens_per_s": S / (ms * 1e-3), "tflops_per_gpu": tf,
                    "frac_of_measured_cublas_sustained": tf / tf_sus, "clocks": clocks(win),
                    "nvlink_bytes_per_rank": (world - 1) * H * S * (D * 2 + 4) if world > 1 else 0,
                    "nvlink_gbs_if_serialised": ((world - 1) * H * S * (D * 2 + 4) / (ms * 1e-3) / 1e9) if world > 1 else 0.0,
                    "note": ("causal with CONTIGUOUS shards: rank 0's keys are visible to every query, so the slowest rank does "
                             "the full (un-halved) work; tflops_per_gpu uses the MEAN work") if causal else "",
                }
            if world > 1 and (S // (2 * world)) % 128 == 0:
                # causal with ZIGZAG shards (rank r owns chunks r and 2W-1-r): same total work, balanced over the ranks.  The K/V
                # values differ from the contiguous runs (same shapes, different order), which does not matter for timing.
                fn = lambda: ta.tree_attention(q, k, v, causal=True, backend="fused", kv_layout="zigzag")
                ms_z, win = _timed(torch, dist, fn, 3, 1, world, dev, barrier)
                flops = 4.0 * S * s_local * D * H * 0.5
                blk["causal_zigzag"] = {"ms": ms_z, "tokens_per_s": S / (ms_z * 1e-3), "tflops_per_gpu": flops / (ms_z * 1e-3) / 1e12,
                                        "speedup_vs_contiguous_causal": blk["causal"]["ms"] / ms_z, "clocks": clocks(win),
                                        "note": "one fused launch per rank over the two-segment shard (kv_seg), combine included"}
            if world > 1:
                # (i) the Sq-sharded output of the reduce-scatter combine (no all-gather of the final tiles), (ii) the same
                # kernel with the cross-GPU combine switched off (local partial only): what the combine costs end to end
                from tree_attention_b200.ops import flash

                fn = lambda: ta.tree_attention(q, k, v, causal=False, backend="fused", output="sharded")
                ms_s, _ = _timed(torch, dist, fn, 3, 1, world, dev, barrier)
                fn = lambda: flash.attention_fwd(q, k, v, D ** -0.5, False, 0, 0)
                ms_l, _ = _timed(torch, dist, fn, 3, 1, world, dev, barrier)
                flops = 4.0 * S * s_local * D * H
                blk["full_sharded_output"] = {"ms": ms_s, "tflops_per_gpu": flops / (ms_s * 1e-3) / 1e12}
                blk["local_partial_only_no_combine"] = {"ms": ms_l, "tflops_per_gpu": flops / (ms_l * 1e-3) / 1e12}
            # correctness spot check: 256 query rows against the fp32 oracle over the gathered sequence
            o = ta.tree_attention(q, k, v, causal=False, backend="fused" if world > 1 else "auto")
            from tree_attention_b200.ops import reference as ref

            rows = slice(S // 2, S // 2 + 256)
            o_p, l_p = ref.attention_partial_ref(q[:, :, rows], k, v, D ** -0.5, False, 0, 0, torch.float32, block=16384)
            if world > 1:
                packed = torch.cat([o_p, l_p[..., None]], -1).contiguous()
                bufs = [torch.empty_like(packed) for _ in range(world)]
                dist.all_gather(bufs, packed)
                o_ref, _ = ref.merge_many([b[..., :-1] for b in bufs], [b[..., -1] for b in bufs])
            else:
                o_ref = o_p
            blk["max_abs_err_vs_oracle_256_rows"] = float((o[:, :, rows].float() - o_ref).abs().max())
            out["fwd_128k_bf16"] = blk
            del q, k, v, o
            torch.cuda.empty_cache()
        except Exception as e:  # noqa: BLE001
            out["fwd_128k_bf16"] = {"error": f"{type(e).__name__}: {e}"[:300]}

    # ------------------------------------------------------------------------------------------ (b) 256K block-scaled fp8 decode
    if want("decode_256k_mxfp8"):
        try:
            from tree_attention_b200.ops.quant import MXFP8SeqTensor, MXFP8Tensor

            S, H, D = 262144 // scale_down, 32, 128
            s_local = S // world
            g = torch.Generator(device=dev).manual_seed(78)
            q = torch.randn(1, H, 1, D, device=dev, generator=g).to(torch.bfloat16)
            nbuf = max(1, -(-4 * (126 << 20) // (2 * H * s_local * D)))   # fp8 bytes; rotate > 4 x L2
            nbuf = min(nbuf, 8)
            caches = []
            for i in range(nbuf):
                _, k, v = ta.make_data((1, H, s_local, D), rank, dev, dtype=torch.bfloat16, seed=200 + i, log=False)
                caches.append((MXFP8Tensor.from_float(k), MXFP8SeqTensor.from_float(v), k if i == 0 else None, v if i == 0 else None))
                if i:
                    del k, v
            it = [0]

            def step():
                kq, vs = caches[it[0] % nbuf][:2]
                it[0] += 1
                return ta.tree_attention(q, kq, vs, backend="fused" if world > 1 else "auto")

            ms, win = _timed(torch, dist, step, 200, 10, world, dev, barrier)
            kv_bytes = 2 * H * s_local * D * 1 + 2 * H * s_local * D // 32     # e4m3 + one UE8M0 byte per 32
            o8 = ta.tree_attention(q, caches[0][0], caches[0][1], backend="fused" if world > 1 else "auto")
            o16 = ta.tree_attention(q, caches[0][2], caches[0][3], backend="fused" if world > 1 else "auto")
            out["decode_256k_mxfp8"] = {
                "seq": S, "heads": H, "head_dim": D, "kv_format": "mxfp8 (e4m3 + UE8M0 per 32): K blocked along channels, V along keys",
                "kernel": "decode_swap_kernel<MX>: tcgen05.mma.kind::mxf8f6f4.block_scale for both GEMMs, fused tree combine",
                "ms": ms, "kv_tokens_per_s": S / (ms * 1e-3), "kv_tokens_per_rank": s_local,
                "hbm_gbs_per_gpu": kv_bytes / (ms * 1e-3) / 1e9, "hbm_frac_of_measured": kv_bytes / (ms * 1e-3) / 1e9 / hbm,
                "max_abs_diff_vs_bf16_kv": float((o8.float() - o16.float()).abs().max()),
                "l2": f"{nbuf} caches x {kv_bytes >> 20} MiB rotated per step", "clocks": clocks(win),
            }
            del caches, q
            torch.cuda.empty_cache()
        except Exception as e:  # noqa: BLE001
            out["decode_256k_mxfp8"] = {"error": f"{type(e).__name__}: {e}"[:300]}

    # ------------------------------------------------------------------------------------------ (c) 1M GQA forward + backward
    if want("gqa_1m_fwd_bwd"):
        try:
            from tree_attention_b200.ops.autograd import tree_attention_func
            from tree_attention_b200.parallel.tree import allreduce_sum

            S, H, HKV, D = (1 << 20) // scale_down, 32, 8, 128
            s_local = S // world
            g = torch.Generator(device=dev).manual_seed(79)
            q = torch.randn(1, H, S, D, device=dev, generator=g).to(torch.bfloat16).requires_grad_(True)
            do = torch.randn(1, H, S, D, device=dev, generator=g).to(torch.bfloat16)
            _, k, v = ta.make_data((1, H, s_local, D), rank, dev, dtype=torch.bfloat16, num_kv_heads=HKV, seed=9, log=False)
            k.requires_grad_(True)
            v.requires_grad_(True)
            res = {}

            def fwd_bwd():
                q.grad = k.grad = v.grad = None
                t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True); t2 = torch.cuda.Event(enable_timing=True)
                t0.record()
                o = tree_attention_func(q, k, v, causal=False, backend="fused" if world > 1 else "auto")
                t1.record()
                o.backward(do)
                t2.record()
                torch.cuda.synchronize()
                res["fwd_ms"], res["bwd_ms"] = t0.elapsed_time(t1), t1.elapsed_time(t2)
                del o

            ms, win = _timed(torch, dist, fwd_bwd, 1, 1, world, dev, barrier)
            parts = torch.tensor([res["fwd_ms"], res["bwd_ms"]], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(parts, op=dist.ReduceOp.MAX)
            fwd_ms, bwd_ms = float(parts[0]), float(parts[1])
            f_fwd = 4.0 * S * s_local * D * H
            blk = {
                "seq": S, "heads": H, "kv_heads": HKV, "head_dim": D, "dtype": "bf16", "causal": False, "kv_tokens_per_rank": s_local,
                "fwd_bwd_ms": ms, "fwd_ms": fwd_ms, "bwd_ms": bwd_ms, "tokens_per_s_fwd_bwd": S / (ms * 1e-3),
                "fwd_tflops_per_gpu": f_fwd / (fwd_ms * 1e-3) / 1e12, "bwd_tflops_per_gpu_5gemm": 2.5 * f_fwd / (bwd_ms * 1e-3) / 1e12,
                "fwd_frac_of_measured_cublas_sustained": f_fwd / (fwd_ms * 1e-3) / 1e12 / tf_sus,
                "bwd_frac_of_measured_cublas_sustained": 2.5 * f_fwd / (bwd_ms * 1e-3) / 1e12 / tf_sus,
                "dq_reduce": "fp32 dQ partials summed over ranks by csrc/reduce.cu (pull reduce-scatter + push all-gather over symmetric memory)",
                "clocks": clocks(win),
            }
            q.grad = k.grad = v.grad = None
            del do
            torch.cuda.empty_cache()
            if world > 1:   # the backward's reduce on its own: 1 GiB of fp32 per rank
                x = torch.randn((1 << 28) // scale_down, device=dev, dtype=torch.float32)
                t_own, _ = _timed(torch, dist, lambda: allreduce_sum(x), 5, 2, world, dev, barrier)
                y = x.clone()
                t_nccl, _ = _timed(torch, dist, lambda: dist.all_reduce(y), 5, 2, world, dev, barrier)
                nbytes = x.numel() * 4
                bus = lambda ms_: 2 * (world - 1) / world * nbytes / (ms_ * 1e-3) / 1e9
                blk["allreduce_sum_1GiB_fp32"] = {"own_symm_ms": t_own, "nccl_ms": t_nccl, "own_busbw_gbs": bus(t_own),
                                                  "nccl_busbw_gbs": bus(t_nccl), "own_over_nccl": t_nccl / t_own}
                del x, y
            out["gqa_1m_fwd_bwd"] = blk
            del q, k, v
            torch.cuda.empty_cache()
        except Exception as e:  # noqa: BLE001
            out["gqa_1m_fwd_bwd"] = {"error": f"{type(e).__name__}: {e}"[:300]}

    if own_sampler and sampler is not None:
        sampler.stop()
    return out


def main():
    import argparse

    import torch
    import torch.distributed as dist

    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/configs.json")
    ap.add_argument("--only", nargs="*", default=None)
    ap.add_argument("--scale-down", type=int, default=1, help="divide every sequence length (smoke-testing the runner on fewer GPUs)")
    a = ap.parse_args()
    import tree_attention_b200 as ta

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    ta.setup(rank, world, local_rank=int(os.environ.get("LOCAL_RANK", str(rank))))
    dev = torch.device("cuda", torch.cuda.current_device())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    res = run_all(ta, world, rank, dev, barrier, ROOT, only=a.only, scale_down=a.scale_down)
    if rank == 0:
        os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
        with open(a.out, "w") as f:
            json.dump({"world": world, "scale_down": a.scale_down, **res}, f, indent=1)
        print(json.dumps({"world": world, "scale_down": a.scale_down, **res}))
    ta.cleanup()


if __name__ == "__main__":
    main()
