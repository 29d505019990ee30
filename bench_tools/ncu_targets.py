"""One kernel per invocation, a few launches, for `ncu --set full -k regex:<kernel>` (bench_tools/ncu_all.sh).
usage: python bench_tools/ncu_targets.py <decode_simt|decode_simt_shard|decode_swap|decode_swap_mx|decode_tc|fwd|fwd_causal|bwd|quant>"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tree_attention_b200.ops import flash, local as L, quant

t = sys.argv[1]
g = torch.Generator(device="cuda").manual_seed(0)
def rnd(*s): return torch.randn(*s, device="cuda", generator=g).bfloat16()
reps = 4
if t in ("decode_simt", "decode_simt_shard"):
    S = 131072 if t == "decode_simt" else 16384          # the headline shape on 1 GPU / its per-rank shard on 8 GPUs
    q, k, v = rnd(1, 32, 1, 128), rnd(1, 32, S, 128), rnd(1, 32, S, 128)
    for _ in range(reps): L.decode_attention(q, k, v, 0.088, impl="simt")
elif t in ("decode_swap", "decode_tc"):
    q, k, v = rnd(1, 32, 1, 128), rnd(1, 8, 131072, 128), rnd(1, 8, 131072, 128)   # GQA 32q/8kv, the 1M config's per-rank shard
    for _ in range(reps): L.decode_attention(q, k, v, 0.088, impl="swap" if t == "decode_swap" else "tc")
elif t == "decode_swap_mx":
    q, k, v = rnd(1, 32, 1, 128), rnd(1, 32, 32768, 128), rnd(1, 32, 32768, 128)    # 256K fp8 config's per-rank shard
    kq, vs = quant.MXFP8Tensor.from_float(k), quant.MXFP8SeqTensor.from_float(v)
    for _ in range(reps): L.decode_attention_mx_tc(q, kq, vs, 0.088)
elif t in ("fwd", "fwd_causal", "bwd"):
    q, k, v = rnd(1, 32, 16384, 128), rnd(1, 32, 16384, 128), rnd(1, 32, 16384, 128)
    causal = t != "fwd"
    for _ in range(reps): o, lse = flash.attention_fwd(q, k, v, 0.088, causal, 0, 0)
    if t == "bwd":
        do = torch.randn_like(q)
        for _ in range(reps): flash.attention_bwd(q, k, v, o, lse, do, 0.088, causal, 0, 0)
elif t == "quant":
    x = rnd(1, 32, 32768, 128)
    for _ in range(reps): quant.MXFP8Tensor.from_float(x); quant.MXFP8SeqTensor.from_float(x)
else:
    raise SystemExit("unknown target")
torch.cuda.synchronize()
print("ok", t)
