"""API-surface parity with /root/reference/model.py and one regression test per
reference defect."""
import inspect
import math

import pytest
import torch

import tree_attention_b200 as ta
from tree_attention_b200.ops import reference as ref


def test_public_names_and_signatures():
    for name in ("setup", "cleanup", "make_data", "flash_res_lse", "tree_decode", "tree_attention"):
        assert hasattr(ta, name), name
    assert list(inspect.signature(ta.setup).parameters)[:2] == ["rank", "world_size"]
    assert list(inspect.signature(ta.make_data).parameters)[:3] == ["shape", "rank", "device"]
    p = inspect.signature(ta.flash_res_lse).parameters
    assert list(p) == ["q", "k", "v", "softmax_scale", "is_causal"]
    assert p["softmax_scale"].default == 1.0 and p["is_causal"].default is False  # model.py:60
    assert list(inspect.signature(ta.tree_decode).parameters)[:6] == ["q", "k", "v", "rank", "world_size", "device"]
    import model  # the root CLI module re-exports the reference's names

    for name in ("setup", "cleanup", "make_data", "flash_res_lse", "tree_decode", "main"):
        assert hasattr(model, name)
    assert list(inspect.signature(model.main).parameters) == ["rank", "world_size"]


def test_D1_make_data_layout_is_bhsd_and_bshd_views():
    q, k, v = ta.make_data((2, 4, 10, 8), 0, "cpu", dtype=torch.float32, log=False)
    assert q.shape == (2, 4, 1, 8) and k.shape == (2, 4, 10, 8) and v.shape == (2, 4, 10, 8)
    q2, k2, v2 = ta.make_data((2, 4, 10, 8), 0, "cpu", dtype=torch.float32, layout="bshd", log=False)
    assert k2.shape == (2, 4, 10, 8) and k2.stride(1) == 8 and k2.stride(2) == 32  # BSHD memory, BHSD view
    o = ta.flash_res_lse(q2, k2, v2)[0]
    assert o.shape == (2, 4, 1, 8)  # attention over the sequence, not over heads


def test_D2_lse_is_logsumexp_of_logits():
    g = torch.Generator().manual_seed(0)
    q = torch.randn(1, 2, 1, 8, generator=g)
    k = torch.randn(1, 2, 16, 8, generator=g)
    v = torch.randn(1, 2, 16, 8, generator=g)
    res, lse = ta.flash_res_lse(q, k, v, softmax_scale=0.5)
    s = (q @ k.transpose(-2, -1)) * 0.5
    assert torch.allclose(lse, torch.logsumexp(s, -1), atol=1e-5)
    assert torch.allclose(res, torch.softmax(s, -1) @ v, atol=1e-5)
    assert lse.shape == res.shape[:-1]


def test_D4_query_is_replicated_kv_is_per_rank():
    q0, k0, _ = ta.make_data((1, 2, 8, 4), 0, "cpu", dtype=torch.float32, log=False)
    q1, k1, _ = ta.make_data((1, 2, 8, 4), 1, "cpu", dtype=torch.float32, log=False)
    assert torch.equal(q0, q1)
    assert not torch.equal(k0, k1)


def test_D5_world1_is_the_local_kernel():
    q, k, v = ta.make_data((1, 2, 33, 8), 0, "cpu", dtype=torch.float32, log=False)
    out = ta.tree_decode(q, k, v, 0, 1, torch.device("cpu"))
    assert torch.allclose(out, ta.flash_res_lse(q, k, v)[0])


def test_D6_causal_is_minus_inf_with_global_offsets():
    g = torch.Generator().manual_seed(1)
    q = torch.randn(1, 1, 4, 8, generator=g)
    k = torch.randn(1, 1, 12, 8, generator=g)
    v = torch.randn(1, 1, 12, 8, generator=g)
    res, _ = ta.flash_res_lse(q, k, v, is_causal=True)
    s = (q @ k.transpose(-2, -1))
    mask = torch.ones(4, 12).tril(diagonal=8).bool()
    exp = torch.softmax(s.masked_fill(~mask, float("-inf")), -1) @ v
    assert torch.allclose(res, exp, atol=1e-5)


def test_D7_scale_defaults():
    g = torch.Generator().manual_seed(2)
    q = torch.randn(1, 1, 1, 16, generator=g)
    k = torch.randn(1, 1, 8, 16, generator=g)
    v = torch.randn(1, 1, 8, 16, generator=g)
    unscaled = torch.softmax(q @ k.transpose(-2, -1), -1) @ v
    scaled = torch.softmax(q @ k.transpose(-2, -1) / math.sqrt(16), -1) @ v
    assert torch.allclose(ta.tree_decode(q, k, v, 0, 1, "cpu"), unscaled, atol=1e-5)   # reference default 1.0
    assert torch.allclose(ta.tree_attention(q, k, v), scaled, atol=1e-5)               # new API: 1/sqrt(d)


def test_D9_fp32_statistics_survive_large_logits():
    # fp16 exp/sub on unscaled d=128 logits overflows in the reference; statistics here are fp32
    g = torch.Generator().manual_seed(3)
    q = (torch.randn(1, 1, 1, 128, generator=g) * 4).half()
    k = (torch.randn(1, 1, 64, 128, generator=g) * 4).half()
    v = torch.randn(1, 1, 64, 128, generator=g).half()
    out, lse = ta.tree_attention(q, k, v, softmax_scale=1.0, return_lse=True)
    assert torch.isfinite(out).all() and torch.isfinite(lse).all()
    assert lse.dtype == torch.float32


def test_tree_decode_requires_group_for_world_gt_1():
    q, k, v = ta.make_data((1, 1, 4, 8), 0, "cpu", dtype=torch.float32, log=False)
    with pytest.raises(RuntimeError):
        ta.tree_decode(q, k, v, 0, 2, "cpu")


def test_gqa_and_multi_query_rows():
    g = torch.Generator().manual_seed(4)
    q = torch.randn(2, 8, 3, 16, generator=g)
    k = torch.randn(2, 2, 40, 16, generator=g)
    v = torch.randn(2, 2, 40, 16, generator=g)
    out, lse = ta.tree_attention(q, k, v, causal=True, return_lse=True)
    o_ref, l_ref = ref.attention_ref(q, k, v, causal=True)
    assert torch.allclose(out.double(), o_ref, atol=1e-5) and torch.allclose(lse.double(), l_ref, atol=1e-5)


def test_bshd_layout_argument():
    g = torch.Generator().manual_seed(5)
    q = torch.randn(1, 2, 4, 8, generator=g)   # (B, S, H, D)
    k = torch.randn(1, 20, 4, 8, generator=g)
    v = torch.randn(1, 20, 4, 8, generator=g)
    out = ta.tree_attention(q, k, v, layout="bshd")
    exp = ta.tree_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)).transpose(1, 2)
    assert out.shape == q.shape and torch.allclose(out, exp)


def test_config_defaults_are_reference_literals():
    from tree_attention_b200.utils.config import TreeAttentionConfig, from_args

    c = TreeAttentionConfig()
    assert (c.batch, c.num_heads, c.seq_len, c.head_dim, c.dtype) == (1, 16, 64000, 128, "fp16")  # model.py:140-145
    assert c.master_port == 12355 and c.log_file == "tree_attention_log.log"                      # model.py:21,160
    c2 = from_args(["--seq-len", "128", "--num-kv-heads", "4", "--no-check"])
    assert c2.seq_len == 128 and c2.kv_heads == 4 and c2.check is False


def test_zigzag_layout_helpers_and_single_rank_passthrough():
    """zigzag_shard / zigzag_unshard are inverses; with one rank (or without a causal mask) kv_layout='zigzag' is the
    contiguous layout; contradictory arguments are rejected."""
    import pytest
    import tree_attention_b200 as ta
    from tree_attention_b200.parallel.tree import zigzag_chunks

    x = torch.arange(2 * 3 * 24 * 4, dtype=torch.float32).view(2, 3, 24, 4)
    for world in (1, 2, 3, 4):
        shards = [ta.zigzag_shard(x, r, world) for r in range(world)]
        assert all(s.shape[2] == 24 // world for s in shards)
        assert torch.equal(ta.zigzag_unshard(shards), x)
        owned = sorted(c for r in range(world) for c in zigzag_chunks(r, world))
        assert owned == list(range(2 * world))           # every chunk owned exactly once
    y = torch.arange(2 * 24 * 3 * 4, dtype=torch.float32).view(2, 24, 3, 4)       # (B, S, H, D): sequence along dim 1
    assert torch.equal(ta.zigzag_unshard([ta.zigzag_shard(y, r, 2, dim=1) for r in range(2)], dim=1), y)
    with pytest.raises(ValueError):
        ta.zigzag_shard(x, 0, 5)                          # 24 % 10 != 0
    g = torch.Generator().manual_seed(0)
    q = torch.randn(1, 2, 8, 16, generator=g)
    k = torch.randn(1, 2, 8, 16, generator=g)
    v = torch.randn(1, 2, 8, 16, generator=g)
    a = ta.tree_attention(q, k, v, causal=True, kv_layout="zigzag")               # world 1: chunks 0 and 1 are adjacent
    b = ta.tree_attention(q, k, v, causal=True)
    assert torch.equal(a, b)
    a = ta.tree_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), causal=True, kv_layout="zigzag", layout="bshd")
    assert torch.allclose(a.transpose(1, 2), b)
    with pytest.raises(ValueError):
        ta.tree_attention(q, k, v, causal=True, kv_layout="ring")
