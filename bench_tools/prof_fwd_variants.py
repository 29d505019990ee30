import os, sys, torch
sys.path.insert(0, os.getcwd())
from tree_attention_b200.ops import flash
g = torch.Generator(device="cuda").manual_seed(0)
q = torch.randn(1, 32, 16384, 128, device="cuda", generator=g).bfloat16()
k = torch.randn(1, 32, 16384, 128, device="cuda", generator=g).bfloat16()
v = torch.randn(1, 32, 16384, 128, device="cuda", generator=g).bfloat16()
for var in (1, 6):
    for _ in range(2):
        flash.attention_fwd(q, k, v, 0.088, False, 0, 0, variant=var)
torch.cuda.synchronize()
