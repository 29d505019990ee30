"""tcgen05 backward (csrc/attn_bwd_sm100.cu) vs the fp32 PyTorch reference backward of the same op."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from tree_attention_b200.ops import flash
from tree_attention_b200.ops import reference as ref

CASES = [
    # b, hq, hkv, sq, s, d, dtype, causal, q_pos0 (None = last), kv_pos0
    (1, 1, 1, 128, 128, 128, torch.bfloat16, False, None, 0),
    (1, 2, 2, 256, 512, 128, torch.bfloat16, False, None, 0),
    (2, 4, 2, 320, 700, 128, torch.bfloat16, True, None, 0),      # GQA, ragged, causal
    (1, 4, 4, 200, 333, 64, torch.float16, True, None, 0),
    (1, 8, 2, 512, 1024, 64, torch.bfloat16, True, None, 0),
    (1, 4, 4, 256, 512, 128, torch.bfloat16, True, 768, 512),     # middle shard of a longer sequence
    (1, 2, 2, 128, 256, 128, torch.bfloat16, True, 50, 1024),     # shard entirely in the future: zero grads
    (1, 16, 4, 1024, 2048, 128, torch.bfloat16, True, None, 0),
]


@pytest.mark.parametrize("case", CASES, ids=[str(i) for i in range(len(CASES))])
def test_bwd_matches_reference(case):
    b, hq, hkv, sq, s, d, dtype, causal, q_pos0, kv_pos0 = case
    g = torch.Generator(device="cuda").manual_seed(1)
    q = torch.randn(b, hq, sq, d, device="cuda", generator=g).to(dtype)
    k = torch.randn(b, hkv, s, d, device="cuda", generator=g).to(dtype)
    v = torch.randn(b, hkv, s, d, device="cuda", generator=g).to(dtype)
    do = torch.randn(b, hq, sq, d, device="cuda", generator=g).to(dtype)
    scale = d ** -0.5
    if q_pos0 is None:
        q_pos0 = s - sq
    o, lse = ref.attention_partial_ref(q, k, v, scale, causal, q_pos0, kv_pos0)
    o = o.to(dtype)
    dq_ref, dk_ref, dv_ref = ref.attention_bwd_ref(q, k, v, o, lse, do, scale, causal, q_pos0, kv_pos0)
    dq, dk, dv = flash.attention_bwd(q, k, v, o, lse, do, scale, causal, q_pos0, kv_pos0)
    torch.cuda.synchronize()
    for name, got, exp in (("dq", dq, dq_ref), ("dk", dk, dk_ref), ("dv", dv, dv_ref)):
        assert not torch.isnan(got).any(), name
        denom = exp.abs().max().item() + 1e-6
        err = (got.float() - exp).abs().max().item() / denom
        assert err < 3e-2, (name, err)


def test_autograd_function_on_gpu():
    from tree_attention_b200.ops.autograd import tree_attention_func

    g = torch.Generator(device="cuda").manual_seed(2)
    q = torch.randn(1, 8, 384, 128, device="cuda", generator=g).bfloat16().requires_grad_(True)
    k = torch.randn(1, 2, 640, 128, device="cuda", generator=g).bfloat16().requires_grad_(True)
    v = torch.randn(1, 2, 640, 128, device="cuda", generator=g).bfloat16().requires_grad_(True)
    do = torch.randn(1, 8, 384, 128, device="cuda", generator=g).bfloat16()
    o = tree_attention_func(q, k, v, causal=True)
    o.backward(do)
    qf, kf, vf = (t.detach().float().requires_grad_(True) for t in (q, k, v))
    o_ref, _ = ref.attention_partial_ref(qf, kf, vf, None, True, 640 - 384, 0)
    o_ref.backward(do.float())
    assert (o.float() - o_ref).abs().max().item() < 2e-2
    for got, exp in ((q.grad, qf.grad), (k.grad, kf.grad), (v.grad, vf.grad)):
        assert (got.float() - exp).abs().max().item() / (exp.abs().max().item() + 1e-6) < 3e-2
