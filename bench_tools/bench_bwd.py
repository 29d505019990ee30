"""Single-GPU throughput of the tcgen05 backward vs FA2 / SDPA backward."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tree_attention_b200.ops import flash
from tree_attention_b200.utils.timing import time_cuda

ap = argparse.ArgumentParser()
ap.add_argument("--seq", type=int, nargs="*", default=[4096, 16384])
ap.add_argument("--heads", type=int, default=32)
ap.add_argument("--kv-heads", type=int, default=None)
ap.add_argument("--dim", type=int, default=128)
ap.add_argument("--steps", type=int, default=5)
a = ap.parse_args()
hkv = a.kv_heads or a.heads
for s in a.seq:
    for causal in (0, 1):
        g = torch.Generator(device="cuda").manual_seed(0)
        q, do = (torch.randn(1, a.heads, s, a.dim, device="cuda", generator=g).bfloat16() for _ in range(2))
        k, v = (torch.randn(1, hkv, s, a.dim, device="cuda", generator=g).bfloat16() for _ in range(2))
        scale = a.dim ** -0.5
        o, lse = flash.attention_fwd(q, k, v, scale, bool(causal), 0, 0)
        flops = 10.0 * s * s * a.dim * a.heads * (0.5 if causal else 1.0)  # 5 GEMMs of the minimal backward
        t = time_cuda(lambda: flash.attention_bwd(q, k, v, o, lse, do, scale, bool(causal), 0, 0), a.steps, 2)
        line = {"seq": s, "causal": causal, "heads": a.heads, "kv_heads": hkv, "own_bwd_ms": round(t["median_ms"], 3),
                "own_bwd_tflops_5gemm": round(flops / t["median_ms"] / 1e9, 1)}
        try:
            from flash_attn import flash_attn_func
            qq, kk, vv = (x.transpose(1, 2).contiguous().requires_grad_(True) for x in (q, k, v))
            oo = flash_attn_func(qq, kk, vv, causal=bool(causal))
            dd = do.transpose(1, 2).contiguous()
            t = time_cuda(lambda: torch.autograd.grad(oo, (qq, kk, vv), dd, retain_graph=True), a.steps, 2)
            line["fa2_bwd_ms"] = round(t["median_ms"], 3)
            line["fa2_bwd_tflops"] = round(flops / t["median_ms"] / 1e9, 1)
        except Exception as e:
            line["fa2_err"] = str(e)[:80]
        try:
            import torch.nn.functional as F
            qq, kk, vv = (x.clone().requires_grad_(True) for x in (q, k, v))
            oo = F.scaled_dot_product_attention(qq, kk, vv, is_causal=bool(causal), enable_gqa=hkv != a.heads)
            t = time_cuda(lambda: torch.autograd.grad(oo, (qq, kk, vv), do, retain_graph=True), a.steps, 2)
            line["sdpa_bwd_ms"] = round(t["median_ms"], 3)
            line["sdpa_bwd_tflops"] = round(flops / t["median_ms"] / 1e9, 1)
        except Exception as e:
            line["sdpa_err"] = str(e)[:80]
        print(json.dumps(line), flush=True)
