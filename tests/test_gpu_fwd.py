"""tcgen05 flash forward (csrc/attn_fwd_sm100.cu) vs the fp32 PyTorch oracle of the same op."""
import pytest
import torch

pytestmark = pytest.mark.gpu

import tree_attention_b200 as ta
from tree_attention_b200.ops import flash
from tree_attention_b200.ops import reference as ref


def _mk(b, hq, hkv, sq, s, d, dtype, seed=0, bshd=False):
    g = torch.Generator(device="cuda").manual_seed(seed)
    if bshd:
        q = torch.randn(b, sq, hq, d, device="cuda", generator=g).to(dtype).transpose(1, 2)
        k = torch.randn(b, s, hkv, d, device="cuda", generator=g).to(dtype).transpose(1, 2)
        v = torch.randn(b, s, hkv, d, device="cuda", generator=g).to(dtype).transpose(1, 2)
    else:
        q = torch.randn(b, hq, sq, d, device="cuda", generator=g).to(dtype)
        k = torch.randn(b, hkv, s, d, device="cuda", generator=g).to(dtype)
        v = torch.randn(b, hkv, s, d, device="cuda", generator=g).to(dtype)
    return q, k, v


CASES = [
    # b, hq, hkv, sq, s, d, dtype, causal, q_pos0 (None = last Sq positions), kv_pos0, bshd
    (1, 1, 1, 128, 128, 128, torch.bfloat16, False, None, 0, False),     # one tile
    (1, 2, 2, 128, 256, 128, torch.bfloat16, False, None, 0, False),     # two kv tiles
    (1, 2, 2, 256, 1024, 128, torch.bfloat16, False, None, 0, False),
    (2, 4, 2, 384, 1000, 128, torch.bfloat16, True, None, 0, False),     # GQA, causal, ragged kv
    (1, 4, 4, 200, 333, 128, torch.float16, True, None, 0, False),       # ragged q and kv, fp16
    (1, 8, 2, 512, 2048, 64, torch.bfloat16, True, None, 0, False),      # head_dim 64
    (2, 2, 2, 100, 700, 64, torch.float16, False, None, 0, False),
    (1, 4, 4, 512, 512, 128, torch.bfloat16, True, 1024, 512, False),    # a middle shard of a longer sequence
    (1, 4, 4, 256, 512, 128, torch.bfloat16, True, 100, 4096, False),    # shard entirely in the future: identity
    (1, 4, 2, 300, 900, 128, torch.bfloat16, True, None, 0, True),       # BSHD-strided
    (1, 2, 2, 1, 4096, 128, torch.bfloat16, False, None, 0, False),      # Sq = 1 through the tensor-core path
    (1, 32, 8, 2048, 4096, 128, torch.bfloat16, True, None, 0, False),   # bigger, GQA 32q/8kv
]


@pytest.mark.parametrize("variant", [1, 6])
@pytest.mark.parametrize("case", CASES, ids=[str(i) for i in range(len(CASES))])
def test_fwd_matches_oracle(case, variant):
    b, hq, hkv, sq, s, d, dtype, causal, q_pos0, kv_pos0, bshd = case
    q, k, v = _mk(b, hq, hkv, sq, s, d, dtype, bshd=bshd)
    scale = d ** -0.5
    if q_pos0 is None:
        q_pos0 = s - sq
    out, lse = flash.attention_fwd(q, k, v, scale, causal, q_pos0, kv_pos0, variant=variant)
    torch.cuda.synchronize()
    o_ref, l_ref = ref.attention_partial_ref(q, k, v, scale, causal, q_pos0, kv_pos0, torch.float32)
    assert not torch.isnan(out).any()
    tol = 2e-2 if dtype == torch.bfloat16 else 5e-3
    err = (out.float() - o_ref).abs().max().item()
    assert err < tol, err
    dead = torch.isinf(l_ref)
    assert torch.equal(torch.isinf(lse) & (lse < 0), dead)
    assert (lse[~dead] - l_ref[~dead]).abs().max().item() < 5e-3 if (~dead).any() else True


def test_large_logits_lazy_rescale():
    # growing maxima along the kv axis force the rescale path on many tiles
    g = torch.Generator(device="cuda").manual_seed(1)
    q = torch.randn(1, 2, 256, 128, device="cuda", generator=g).bfloat16()
    k = torch.randn(1, 2, 2048, 128, device="cuda", generator=g)
    k = (k * torch.linspace(0.2, 6.0, 2048, device="cuda")[None, None, :, None]).bfloat16()
    v = torch.randn(1, 2, 2048, 128, device="cuda", generator=g).bfloat16()
    out, lse = flash.attention_fwd(q, k, v, 0.3, False, 0, 0)
    o_ref, l_ref = ref.attention_partial_ref(q, k, v, 0.3)
    assert (out.float() - o_ref).abs().max().item() < 3e-2
    assert (lse - l_ref).abs().max().item() < 2e-2


def test_public_api_routes_prefill_to_tcgen05():
    q, k, v = _mk(1, 8, 8, 1024, 1024, 128, torch.bfloat16, seed=2)
    out, lse = ta.tree_attention(q, k, v, causal=True, return_lse=True)
    o_ref, l_ref = ref.attention_partial_ref(q, k, v, None, True, 0, 0)
    assert (out.float() - o_ref).abs().max().item() < 2e-2


SEG_CASES = [
    # hq, hkv, sq, s, seg_len, d, q_pos0, kv_pos0, seg_gap
    (2, 2, 1024, 512, 256, 128, 0, 0, 512),        # zigzag shard of rank 0 of 2 over a 1024-long sequence, full query block
    (2, 2, 1024, 512, 256, 128, 0, 256, 0),        # rank 1 of 2: chunks 1 and 2 are adjacent (gap 0 = contiguous)
    (4, 2, 700, 768, 384, 128, 900, 384, 1152),    # GQA, ragged query block in the middle of a longer sequence
    (2, 2, 512, 1024, 512, 64, 300, 0, 2048),      # head_dim 64; second segment entirely in the future for most rows
    (2, 2, 300, 640, 128, 128, 100, 0, 128),       # unequal segments (128 + 512 rows), ragged q
]


@pytest.mark.parametrize("variant", [1, 6])
@pytest.mark.parametrize("case", SEG_CASES, ids=[str(i) for i in range(len(SEG_CASES))])
def test_fwd_two_segment_shard_matches_oracle(case, variant):
    """kv_seg=(seg_len, seg_gap): local rows >= seg_len sit seg_gap positions further on (zigzag sharding of a causal
    sequence).  Oracle: the two segments as separate causal partials at their own positions, merged."""
    hq, hkv, sq, s, seg_len, d, q_pos0, kv_pos0, gap = case
    q, k, v = _mk(1, hq, hkv, sq, s, d, torch.bfloat16, seed=11)
    scale = d ** -0.5
    out, lse = flash.attention_fwd(q, k, v, scale, True, q_pos0, kv_pos0, variant=variant, kv_seg=(seg_len, gap))
    torch.cuda.synchronize()
    pa = ref.attention_partial_ref(q, k[:, :, :seg_len], v[:, :, :seg_len], scale, True, q_pos0, kv_pos0, torch.float32)
    pb = ref.attention_partial_ref(q, k[:, :, seg_len:], v[:, :, seg_len:], scale, True, q_pos0, kv_pos0 + gap + seg_len,
                                   torch.float32)
    o_ref, l_ref = ref.merge_many([pa[0], pb[0]], [pa[1], pb[1]])
    assert not torch.isnan(out).any()
    assert (out.float() - o_ref).abs().max().item() < 2e-2
    dead = torch.isinf(l_ref)
    assert torch.equal(torch.isinf(lse) & (lse < 0), dead)
    if (~dead).any():
        assert (lse[~dead] - l_ref[~dead]).abs().max().item() < 5e-3
    # and the layout is honoured: the one-segment interpretation of the same tensors differs whenever the gap matters
    if gap > 0 and q_pos0 + sq - 1 >= kv_pos0 + seg_len:
        o1, _ = flash.attention_fwd(q, k, v, scale, True, q_pos0, kv_pos0, variant=variant)
        assert (o1.float() - o_ref).abs().max().item() > 1e-3
