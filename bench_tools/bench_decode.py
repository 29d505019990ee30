"""Single-GPU decode latency: bf16 KV vs block-scaled fp8 KV, with achieved HBM GB/s."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tree_attention_b200.ops import local as L
from tree_attention_b200.ops.quant import FP8ChannelTensor, MXFP8SeqTensor, MXFP8Tensor
from tree_attention_b200.utils.timing import time_cuda

ap = argparse.ArgumentParser()
ap.add_argument("--seq", type=int, nargs="*", default=[16384, 131072])
ap.add_argument("--heads", type=int, default=32)
ap.add_argument("--kv-heads", type=int, default=None)
ap.add_argument("--steps", type=int, default=200)
a = ap.parse_args()
hkv = a.kv_heads or a.heads
for s in a.seq:
    g = torch.Generator(device="cuda").manual_seed(0)
    q = torch.randn(1, a.heads, 1, 128, device="cuda", generator=g).bfloat16()
    kvs = [(torch.randn(1, hkv, s, 128, device="cuda", generator=g).bfloat16(), torch.randn(1, hkv, s, 128, device="cuda", generator=g).bfloat16())
           for _ in range(max(1, min(8, (600 << 20) // (4 * hkv * s * 128))))]
    mx = [(MXFP8Tensor.from_float(k), MXFP8Tensor.from_float(v)) for k, v in kvs]
    c8 = [(FP8ChannelTensor.from_float(k), FP8ChannelTensor.from_float(v)) for k, v in kvs]
    mxs = [(MXFP8Tensor.from_float(k), MXFP8SeqTensor.from_float(v)) for k, v in kvs]
    n = len(kvs)
    i = [0]
    def f16():
        i[0] += 1
        return L.decode_attention(q, *kvs[i[0] % n], 0.088, return_lse=False, impl="simt")
    def fc8():
        i[0] += 1
        return L.decode_attention_fp8(q, *c8[i[0] % n], 0.088, return_lse=False, impl="tc")
    def fsw():
        i[0] += 1
        return L.decode_attention(q, *kvs[i[0] % n], 0.088, return_lse=False, impl="swap")
    def fsw8():
        i[0] += 1
        return L.decode_attention_fp8(q, *c8[i[0] % n], 0.088, return_lse=False, impl="swap")
    def ftc():
        i[0] += 1
        return L.decode_attention(q, *kvs[i[0] % n], 0.088, return_lse=False, impl="tc")
    def f8():
        i[0] += 1
        return L.decode_attention_mxfp8(q, *mx[i[0] % n], 0.088, return_lse=False)
    def fmxtc():
        i[0] += 1
        return L.decode_attention_mx_tc(q, *mxs[i[0] % n], 0.088, return_lse=False)
    t16 = time_cuda(f16, a.steps, 10)["median_ms"]
    tmxtc = time_cuda(fmxtc, a.steps, 10)["median_ms"]
    t8 = time_cuda(f8, a.steps, 10)["median_ms"]
    ttc = time_cuda(ftc, a.steps, 10)["median_ms"]
    tc8 = time_cuda(fc8, a.steps, 10)["median_ms"]
    bc8 = 2 * hkv * s * 128
    tsw = time_cuda(fsw, a.steps, 10)["median_ms"]
    tsw8 = time_cuda(fsw8, a.steps, 10)["median_ms"]
    b16 = 2 * hkv * s * 128 * 2
    b8 = 2 * hkv * s * (128 + 4)
    print(json.dumps({"seq": s, "heads": a.heads, "kv_heads": hkv, "bf16_us": round(t16 * 1e3, 1), "bf16_gbs": round(b16 / t16 / 1e6, 0),
                      "bf16_tcgen05_us": round(ttc * 1e3, 1), "bf16_tcgen05_gbs": round(b16 / ttc / 1e6, 0),
                      "bf16_swapAB_us": round(tsw * 1e3, 1), "bf16_swapAB_gbs": round(b16 / tsw / 1e6, 0),
                      "fp8_swapAB_us": round(tsw8 * 1e3, 1), "fp8_swapAB_gbs": round(bc8 / tsw8 / 1e6, 0),
                      "fp8_tcgen05_us": round(tc8 * 1e3, 1), "fp8_tcgen05_gbs": round(bc8 / tc8 / 1e6, 0),
                      "mxfp8_block_scaled_tcgen05_us": round(tmxtc * 1e3, 1), "mxfp8_block_scaled_tcgen05_gbs": round(b8 / tmxtc / 1e6, 0),
                      "mxfp8_us": round(t8 * 1e3, 1), "mxfp8_gbs": round(b8 / t8 / 1e6, 0), "speedup": round(t16 / t8, 2)}), flush=True)
