"""Backward of tree attention: dK/dV local, dQ summed over ranks.  CPU / gloo."""
import pytest
import torch

from _dist_utils import run_distributed


def test_single_rank_matches_autograd_of_oracle():
    from tree_attention_b200.ops import reference as ref
    from tree_attention_b200.ops.autograd import tree_attention_func

    g = torch.Generator().manual_seed(0)
    q = torch.randn(2, 4, 5, 16, generator=g, requires_grad=True)
    k = torch.randn(2, 2, 30, 16, generator=g, requires_grad=True)
    v = torch.randn(2, 2, 30, 16, generator=g, requires_grad=True)
    do = torch.randn(2, 4, 5, 16, generator=g)
    o = tree_attention_func(q, k, v, causal=True)
    o.backward(do)
    got = [t.grad.clone() for t in (q, k, v)]
    for t in (q, k, v):
        t.grad = None
    o2, _ = ref.attention_partial_ref(q, k, v, causal=True, q_pos0=25)
    o2.backward(do)
    for a, b in zip(got, (q.grad, k.grad, v.grad)):
        assert torch.allclose(a, b, atol=1e-5), (a - b).abs().max()


def _worker(rank, world):
    from tree_attention_b200.ops import reference as ref
    from tree_attention_b200.ops.autograd import tree_attention_func

    g = torch.Generator().manual_seed(1)
    s_local = 12
    q = torch.randn(1, 4, 6, 8, generator=g)
    kf = torch.randn(1, 2, s_local * world, 8, generator=g)
    vf = torch.randn(1, 2, s_local * world, 8, generator=g)
    do = torch.randn(1, 4, 6, 8, generator=g)
    # oracle on the full sequence
    qo, ko, vo = (t.clone().requires_grad_(True) for t in (q, kf, vf))
    o_ref, _ = ref.attention_partial_ref(qo, ko, vo, causal=True, q_pos0=s_local * world - 6)
    o_ref.backward(do)
    sl = slice(rank * s_local, (rank + 1) * s_local)
    ql = q.clone().requires_grad_(True)
    kl = kf[:, :, sl].clone().requires_grad_(True)
    vl = vf[:, :, sl].clone().requires_grad_(True)
    o = tree_attention_func(ql, kl, vl, causal=True)
    assert torch.allclose(o, o_ref.detach(), atol=1e-5)
    o.backward(do)
    assert torch.allclose(ql.grad, qo.grad, atol=1e-5)
    assert torch.allclose(kl.grad, ko.grad[:, :, sl], atol=1e-5)
    assert torch.allclose(vl.grad, vo.grad[:, :, sl], atol=1e-5)


@pytest.mark.parametrize("world", [2, 3])
def test_distributed_backward_gloo(world, port):
    run_distributed(_worker, world, port)


def test_module_grad_path_passes_backend_and_schedule(monkeypatch):
    """TreeAttention.forward must hand its backend / schedule to the differentiable path too."""
    import tree_attention_b200.ops.autograd as ag
    from tree_attention_b200.models.tree_attention import TreeAttention

    seen = {}
    real = ag.tree_attention_func

    def spy(*a, **kw):
        seen.update(kw)
        return real(*a, **kw)

    monkeypatch.setattr(ag, "tree_attention_func", spy)
    m = TreeAttention(causal=True, backend="collective", schedule="butterfly")
    q = torch.randn(1, 2, 3, 16, requires_grad=True)
    k = torch.randn(1, 2, 9, 16)
    v = torch.randn(1, 2, 9, 16)
    m(q, k, v).sum().backward()
    assert seen.get("backend") == "collective" and seen.get("schedule") == "butterfly"
    assert q.grad is not None and torch.isfinite(q.grad).all()
