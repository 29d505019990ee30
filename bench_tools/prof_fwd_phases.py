"""Where does a softmax warp of attn_fwd_kernel spend its cycles?  (clock64 stamps of thread 0 of CTA 0)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tree_attention_b200 import _build
from tree_attention_b200.ops import flash

C = _build.load()
g = torch.Generator(device="cuda").manual_seed(0)
for s in (16384,):
    q = torch.randn(1, 32, s, 128, device="cuda", generator=g).bfloat16()
    k = torch.randn(1, 32, s, 128, device="cuda", generator=g).bfloat16()
    v = torch.randn(1, 32, s, 128, device="cuda", generator=g).bfloat16()
    for variant in (1, 6):
        for _ in range(3):
            flash.attention_fwd(q, k, v, 0.088, False, 0, 0, variant=variant)
        torch.cuda.synchronize()
        w, f, e, st, n = C.attn_fwd_phase_cycles()
        print(f"variant {variant} seq {s}: tiles {n}; per tile cycles: wait S {w / n:.0f}, fast path {f / n:.0f}, exact path {e / n:.0f}, "
              f"P store + signal {st / n:.0f}, total {(w + f + e + st) / n:.0f}")
