"""The (o, lse) combine is an associative, commutative monoid with identity (0, -inf)."""
import itertools

import pytest
import torch
from hypothesis import given, settings, strategies as st

from tree_attention_b200.ops import reference as ref


def _rand_partial(gen, rows=5, d=8, dead_frac=0.0):
    o = torch.randn(rows, d, generator=gen, dtype=torch.float64)
    lse = torch.randn(rows, generator=gen, dtype=torch.float64) * 5
    if dead_frac:
        dead = torch.rand(rows, generator=gen) < dead_frac
        lse[dead] = float("-inf")
        o[dead] = 0
    return o, lse


@settings(max_examples=40, deadline=None)
@given(st.integers(0, 10_000))
def test_associative(seed):
    g = torch.Generator().manual_seed(seed)
    a, b, c = (_rand_partial(g, dead_frac=0.2) for _ in range(3))
    ab_c = ref.merge_pair(*ref.merge_pair(*a, *b), *c)
    a_bc = ref.merge_pair(*a, *ref.merge_pair(*b, *c))
    assert torch.allclose(ab_c[0], a_bc[0], atol=1e-12)
    assert torch.allclose(ab_c[1], a_bc[1], atol=1e-12)


@settings(max_examples=20, deadline=None)
@given(st.integers(0, 10_000))
def test_commutative_and_identity(seed):
    g = torch.Generator().manual_seed(seed)
    a, b = _rand_partial(g), _rand_partial(g)
    ab, ba = ref.merge_pair(*a, *b), ref.merge_pair(*b, *a)
    assert torch.allclose(ab[0], ba[0], atol=1e-13) and torch.allclose(ab[1], ba[1], atol=1e-13)
    ident = (torch.zeros_like(a[0]), torch.full_like(a[1], float("-inf")))
    ai = ref.merge_pair(*a, *ident)
    assert torch.allclose(ai[0], a[0], atol=1e-13) and torch.allclose(ai[1], a[1], atol=1e-13)
    ii = ref.merge_pair(*ident, *ident)
    assert torch.all(ii[0] == 0) and torch.all(torch.isinf(ii[1]) & (ii[1] < 0))
    assert not torch.isnan(ii[0]).any()


def test_permutation_invariance_and_schedules():
    g = torch.Generator().manual_seed(0)
    parts = [_rand_partial(g, dead_frac=0.3) for _ in range(5)]
    base = ref.merge_many([p[0] for p in parts], [p[1] for p in parts])
    for perm in itertools.islice(itertools.permutations(range(5)), 0, 120, 7):
        o, l = ref.merge_many([parts[i][0] for i in perm], [parts[i][1] for i in perm])
        assert torch.allclose(o, base[0], atol=1e-12) and torch.allclose(l, base[1], atol=1e-12)
        o, l = ref.merge_tree([parts[i][0] for i in perm], [parts[i][1] for i in perm])
        assert torch.allclose(o, base[0], atol=1e-12) and torch.allclose(l, base[1], atol=1e-12)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("causal", [False, True])
def test_sharded_equals_monolithic(dtype, world, causal):
    g = torch.Generator().manual_seed(world)
    b, hq, hkv, sq, s, d = 2, 4, 2, 3, 8 * world * 3, 16
    q = torch.randn(b, hq, sq, d, generator=g).to(dtype)
    k = torch.randn(b, hkv, s, d, generator=g).to(dtype)
    v = torch.randn(b, hkv, s, d, generator=g).to(dtype)
    o_ref, l_ref = ref.attention_ref(q, k, v, causal=causal)
    ks, vs = k.chunk(world, dim=2), v.chunk(world, dim=2)
    for sched in ("flat", "tree"):
        o, l = ref.sharded_attention_ref(q, ks, vs, causal=causal, schedule=sched)
        assert torch.allclose(o.double(), o_ref, atol=2e-5), (o.double() - o_ref).abs().max()
        assert torch.allclose(l.double(), l_ref, atol=2e-5)


def test_fully_masked_shard_is_identity():
    # causal decode at an early position: later shards are fully masked and must contribute nothing
    g = torch.Generator().manual_seed(3)
    q = torch.randn(1, 2, 1, 8, generator=g)
    k = torch.randn(1, 2, 32, 8, generator=g)
    v = torch.randn(1, 2, 32, 8, generator=g)
    o_ref, l_ref = ref.attention_partial_ref(q, k[:, :, :8], v[:, :, :8], causal=True, q_pos0=7)
    o, l = ref.sharded_attention_ref(q, k.chunk(4, 2), v.chunk(4, 2), causal=True, q_pos0=7)
    assert torch.allclose(o, o_ref, atol=1e-6) and torch.allclose(l, l_ref, atol=1e-6)
    o_dead, l_dead = ref.attention_partial_ref(q, k[:, :, 8:16], v[:, :, 8:16], causal=True, q_pos0=7, kv_pos0=8)
    assert torch.all(o_dead == 0) and torch.all(torch.isinf(l_dead))


def test_blockwise_partial_matches_dense():
    g = torch.Generator().manual_seed(5)
    q = torch.randn(1, 4, 5, 16, generator=g)
    k = torch.randn(1, 2, 300, 16, generator=g)
    v = torch.randn(1, 2, 300, 16, generator=g)
    for causal in (False, True):
        o1, l1 = ref.attention_partial_ref(q, k, v, causal=causal, q_pos0=295)
        o2, l2 = ref.attention_partial_ref(q, k, v, causal=causal, q_pos0=295, block=64)
        assert torch.allclose(o1, o2, atol=1e-5) and torch.allclose(l1, l2, atol=1e-5)


def test_backward_reference_matches_autograd():
    g = torch.Generator().manual_seed(7)
    q = torch.randn(1, 4, 6, 8, generator=g, dtype=torch.float64, requires_grad=True)
    k = torch.randn(1, 2, 20, 8, generator=g, dtype=torch.float64, requires_grad=True)
    v = torch.randn(1, 2, 20, 8, generator=g, dtype=torch.float64, requires_grad=True)
    o, lse = ref.attention_partial_ref(q, k, v, causal=True, q_pos0=14, compute_dtype=torch.float64)
    do = torch.randn(o.shape, generator=g, dtype=torch.float64)
    o.backward(do)
    dq, dk, dv = ref.attention_bwd_ref(q.detach().float(), k.detach().float(), v.detach().float(), o.detach().float(),
                                       lse.detach().float(), do.float(), causal=True, q_pos0=14)
    assert torch.allclose(dq.double(), q.grad, atol=1e-4)
    assert torch.allclose(dk.double(), k.grad, atol=1e-4)
    assert torch.allclose(dv.double(), v.grad, atol=1e-4)
