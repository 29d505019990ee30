"""torch.ops.tree_attention.partial: registered operator with fake kernel and autograd through BOTH outputs."""
import torch

from tree_attention_b200.ops import library, reference as ref


def _inputs(sq=5, s=37, hq=4, hkv=2, d=16, seed=0):
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(1, hq, sq, d, generator=g, dtype=torch.float64)
    k = torch.randn(1, hkv, s, d, generator=g, dtype=torch.float64)
    v = torch.randn(1, hkv, s, d, generator=g, dtype=torch.float64)
    return q, k, v


def test_operator_matches_oracle_and_has_fake_impl():
    q, k, v = _inputs()
    o, lse = torch.ops.tree_attention.partial(q.float(), k.float(), v.float(), 0.3, True, 32, 0)
    o_ref, l_ref = ref.attention_partial_ref(q.float(), k.float(), v.float(), 0.3, True, 32, 0)
    assert torch.allclose(o, o_ref, atol=1e-5) and torch.allclose(lse, l_ref, atol=1e-5)
    with torch._subclasses.fake_tensor.FakeTensorMode():
        fq = torch.empty(2, 4, 5, 16)
        fk = torch.empty(2, 2, 37, 16)
        fo, fl = torch.ops.tree_attention.partial(fq, fk, fk, 0.3, False, 0, 0)
        assert fo.shape == (2, 4, 5, 16) and fl.shape == (2, 4, 5) and fl.dtype == torch.float32


def test_gradients_flow_through_o_and_lse_and_a_user_written_combine():
    """Two shards combined by hand with the monoid formula == attention over the concatenation, gradients included."""
    q, k, v = _inputs(sq=3, s=40)
    scale = 0.25
    qa = q.float().clone().requires_grad_(True)
    ka, va = k.float().clone().requires_grad_(True), v.float().clone().requires_grad_(True)
    parts = []
    for lo, hi in ((0, 24), (24, 40)):
        parts.append(library.attention_partial_op(qa, ka[:, :, lo:hi], va[:, :, lo:hi], scale, True, 37, lo))
    lse = torch.logsumexp(torch.stack([p[1] for p in parts]), dim=0)
    out = sum(p[0] * torch.exp(p[1] - lse)[..., None] for p in parts)
    loss = (out * torch.linspace(-1, 1, out.numel()).view_as(out)).sum() + lse.sum() * 0.1
    loss.backward()

    qb = q.float().clone().requires_grad_(True)
    kb, vb = k.float().clone().requires_grad_(True), v.float().clone().requires_grad_(True)
    g = qb.shape[1] // kb.shape[1]
    sc = torch.matmul(qb, kb.repeat_interleave(g, 1).transpose(-1, -2)) * scale
    rows = torch.arange(3).view(3, 1) + 37
    cols = torch.arange(40).view(1, 40)
    sc = sc.masked_fill(cols > rows, float("-inf"))
    lse_b = torch.logsumexp(sc, -1)
    out_b = torch.matmul(torch.softmax(sc, -1), vb.repeat_interleave(g, 1))
    loss_b = (out_b * torch.linspace(-1, 1, out_b.numel()).view_as(out_b)).sum() + lse_b.sum() * 0.1
    loss_b.backward()
    assert torch.allclose(out, out_b, atol=1e-5)
    for a, b in ((qa, qb), (ka, kb), (va, vb)):
        assert torch.allclose(a.grad, b.grad, atol=2e-4), (a.grad - b.grad).abs().max()


def test_opcheck():
    q, k, v = _inputs(sq=2, s=9, d=8)
    args = (q.float().requires_grad_(True), k.float().requires_grad_(True), v.float().requires_grad_(True), 0.5, False, 0, 0)
    torch.library.opcheck(torch.ops.tree_attention.partial.default, args,
                          test_utils=("test_schema", "test_faketensor", "test_autograd_registration"))
