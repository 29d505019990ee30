"""Single-GPU throughput of the tcgen05 forward vs library kernels (FA2 = mma.sync recompiled, SDPA/cuDNN)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tree_attention_b200.ops import flash
from tree_attention_b200.utils.timing import time_cuda

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import time


def timed_with_clocks(fn, steps, warmup):
    """time_cuda plus SM clock / throttle sampling while a >= 1.5 s loop of the same call runs."""
    from bench import ClockSampler

    t = time_cuda(fn, steps, warmup)
    smp = ClockSampler(torch.cuda.current_device(), period_ms=50)
    time.sleep(0.2)
    t0 = time.time()
    n = max(steps, int(1500.0 / max(t["median_ms"], 1e-3)))
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    t1 = time.time()
    smp.stop()
    c = smp.summary(t0 + 0.3, t1)
    t["sustained_ms"] = (t1 - t0) * 1e3 / n
    t["clocks"] = c
    return t


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seq", type=int, nargs="*", default=[4096, 16384, 32768])
    ap.add_argument("--heads", type=int, default=32)
    ap.add_argument("--kv-heads", type=int, default=None)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--causal", type=int, nargs="*", default=[0, 1])
    ap.add_argument("--libs", type=int, default=1)
    a = ap.parse_args()
    peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {}
    peak = peaks.get("bf16_tflops", 1590.0)
    hkv = a.kv_heads or a.heads
    for s in a.seq:
        for causal in a.causal:
            g = torch.Generator(device="cuda").manual_seed(0)
            q = torch.randn(1, a.heads, s, a.dim, device="cuda", generator=g).bfloat16()
            k = torch.randn(1, hkv, s, a.dim, device="cuda", generator=g).bfloat16()
            v = torch.randn(1, hkv, s, a.dim, device="cuda", generator=g).bfloat16()
            flops = 4.0 * s * s * a.dim * a.heads * (0.5 if causal else 1.0)
            scale = a.dim ** -0.5
            res = {}
            t = timed_with_clocks(lambda: flash.attention_fwd(q, k, v, scale, bool(causal), 0, 0), a.steps, a.warmup)
            res["tcgen05_own"] = t["median_ms"]
            res["tcgen05_own_sustained"] = t["sustained_ms"]
            own_clocks = t["clocks"]
            t = time_cuda(lambda: flash.attention_fwd(q, k, v, scale, bool(causal), 0, 0, variant=6), a.steps, a.warmup)
            res["tcgen05_own_v6_q_in_tmem"] = t["median_ms"]
            if a.libs:
                try:
                    from flash_attn import flash_attn_func

                    qq, kk, vv = (x.transpose(1, 2).contiguous() for x in (q, k, v))
                    t = time_cuda(lambda: flash_attn_func(qq, kk, vv, causal=bool(causal)), a.steps, a.warmup)
                    res["flash_attn2"] = t["median_ms"]
                except Exception as e:
                    res["flash_attn2_err"] = str(e)[:80]
                try:
                    import torch.nn.functional as F

                    t = timed_with_clocks(lambda: F.scaled_dot_product_attention(q, k, v, is_causal=bool(causal), enable_gqa=hkv != a.heads),
                                          a.steps, a.warmup)
                    res["sdpa"] = t["median_ms"]
                    res["sdpa_sustained"] = t["sustained_ms"]
                    line_sdpa_clocks = t["clocks"]
                except Exception as e:
                    res["sdpa_err"] = str(e)[:80]
            line = {"seq": s, "causal": causal, "heads": a.heads, "kv_heads": hkv, "dim": a.dim}
            for kname, ms in res.items():
                if isinstance(ms, float):
                    line[kname + "_ms"] = round(ms, 4)
                    line[kname + "_tflops"] = round(flops / ms / 1e9, 1)
                else:
                    line[kname] = ms
            line["own_clocks"] = own_clocks
            if "sdpa" in res:
                line["sdpa_clocks"] = line_sdpa_clocks
            line["own_frac_of_measured_cublas_peak"] = round(flops / res["tcgen05_own"] / 1e9 / peak, 3)
            print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
