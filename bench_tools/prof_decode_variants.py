import os, sys, torch
sys.path.insert(0, os.getcwd())
from tree_attention_b200.ops import local as L
from tree_attention_b200.ops.quant import FP8ChannelTensor
g = torch.Generator(device="cuda").manual_seed(0)
q = torch.randn(1, 32, 1, 128, device="cuda", generator=g).bfloat16()
k = torch.randn(1, 32, 131072, 128, device="cuda", generator=g).bfloat16()
v = torch.randn(1, 32, 131072, 128, device="cuda", generator=g).bfloat16()
k8, v8 = FP8ChannelTensor.from_float(k), FP8ChannelTensor.from_float(v)
for _ in range(2):
    L.decode_attention(q, k, v, 0.088, impl="swap")
    L.decode_attention_fp8(q, k8, v8, 0.088, impl="swap")
torch.cuda.synchronize()
