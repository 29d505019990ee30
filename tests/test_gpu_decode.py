"""Streaming decode kernel (csrc/decode_simt.cu) vs the fp32 PyTorch oracle of the same op."""
import pytest
import torch

pytestmark = pytest.mark.gpu

import tree_attention_b200 as ta
from tree_attention_b200.ops import local as L
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
    # b, hq, hkv, sq, s, d, dtype, causal, bshd
    (1, 1, 1, 1, 1024, 128, torch.bfloat16, False, False),   # BASELINE config #1 shape class (single head)
    (1, 16, 16, 1, 300, 128, torch.float16, False, False),    # ragged tail, reference dtype
    (2, 8, 8, 1, 4096, 128, torch.bfloat16, False, False),
    (1, 32, 32, 1, 16384, 128, torch.bfloat16, False, False),  # north-star per-rank shard at W=8
    (1, 32, 8, 1, 8192, 128, torch.bfloat16, False, False),   # GQA 32q/8kv -> 4 rows
    (2, 8, 2, 2, 2000, 128, torch.bfloat16, True, False),     # 8 rows -> two passes, causal
    (1, 12, 4, 3, 777, 64, torch.float16, True, False),       # head_dim 64, 9 rows -> 3 passes
    (1, 4, 4, 1, 5000, 64, torch.bfloat16, False, False),
    (2, 8, 4, 1, 3000, 128, torch.bfloat16, False, True),     # BSHD-strided inputs
    (1, 2, 2, 1, 100000, 128, torch.bfloat16, False, False),  # few heads, many splits per head
    (3, 5, 5, 1, 129, 128, torch.bfloat16, True, False),
    (1, 32, 8, 8, 3000, 128, torch.bfloat16, True, False),    # 32 packed rows (GQA 4 x 8 tokens: speculative decode)
    (1, 16, 2, 16, 1111, 64, torch.float16, True, False),     # 128 packed rows: a full tile
]


@pytest.mark.parametrize("impl", ["simt", "tc", "swap"])
@pytest.mark.parametrize("case", CASES, ids=[str(i) for i in range(len(CASES))])
def test_decode_matches_oracle(case, impl):
    b, hq, hkv, sq, s, d, dtype, causal, bshd = case
    if impl == "swap" and (d != 128 or (hq // hkv) * sq > 16):
        pytest.skip("swap-AB decode: head_dim 128, <= 16 packed query rows")
    q, k, v = _mk(b, hq, hkv, sq, s, d, dtype, bshd=bshd)
    scale = d ** -0.5
    q_pos0 = s - sq
    out, lse = L.decode_attention(q, k, v, scale, causal, q_pos0, 0, impl=impl)
    o_ref, l_ref = ref.attention_partial_ref(q, k, v, scale, causal, q_pos0, 0, torch.float32, block=16384)
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    tol = 1.5e-2 if dtype == torch.bfloat16 else 4e-3
    assert (out.float() - o_ref).abs().max().item() < tol
    assert (lse - l_ref).abs().max().item() < 2e-3


@pytest.mark.parametrize("impl", ["simt", "tc", "swap"])
def test_causal_offsets_and_fully_masked_shard(impl):
    q, k, v = _mk(1, 4, 4, 1, 1000, 128, torch.bfloat16, seed=3)
    # shard starts AFTER the query position: every key masked -> identity (0, -inf), no NaN
    out, lse = L.decode_attention(q, k, v, 0.1, True, q_pos0=10, kv_pos0=500, impl=impl)
    assert torch.all(out == 0) and torch.all(torch.isinf(lse) & (lse < 0))
    # partially visible shard
    out, lse = L.decode_attention(q, k, v, 0.1, True, q_pos0=700, kv_pos0=500, impl=impl)
    o_ref, l_ref = ref.attention_partial_ref(q, k, v, 0.1, True, 700, 500)
    assert (out.float() - o_ref).abs().max().item() < 1.5e-2 and (lse - l_ref).abs().max().item() < 2e-3


def test_public_api_dispatches_to_kernel_and_matches():
    q, k, v = ta.make_data((1, 16, 64000, 128), 0, "cuda", dtype=torch.float16, log=False)  # model.py:140-145
    out = ta.tree_decode(q, k, v, 0, 1, torch.device("cuda"))       # reference shim: scale 1.0
    o_ref, _ = ref.attention_partial_ref(q, k, v, 1.0, block=16384)
    assert (out.float() - o_ref).abs().max().item() < 1e-2
    res, lse = ta.flash_res_lse(q, k, v)
    assert res.shape == (1, 16, 1, 128) and lse.shape == (1, 16, 1) and lse.dtype == torch.float32


@pytest.mark.parametrize("impl", ["simt", "tc", "swap"])
def test_repeated_calls_are_deterministic_and_reset_tickets(impl):
    q, k, v = _mk(2, 8, 8, 1, 6000, 128, torch.bfloat16, seed=5)
    first, _ = L.decode_attention(q, k, v, 0.088, False, impl=impl)
    first = first.clone()
    for _ in range(200):
        out, _ = L.decode_attention(q, k, v, 0.088, False, impl=impl)
    torch.cuda.synchronize()
    assert torch.equal(out, first)


def test_cuda_graph_replay():
    q, k, v = _mk(1, 8, 8, 1, 4096, 128, torch.bfloat16, seed=7)
    out = torch.empty_like(q)
    lse = torch.empty(1, 8, 1, device="cuda", dtype=torch.float32)
    L.decode_attention(q, k, v, 0.088, out=out, lse=lse)  # warm-up (workspace allocation)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g):
            L.decode_attention(q, k, v, 0.088, out=out, lse=lse)
    ref_out = out.clone()
    for i in range(3):
        q.copy_(torch.randn_like(q))
        g.replay()
        torch.cuda.synchronize()
        exp, _ = ref.attention_partial_ref(q, k, v, 0.088)
        assert (out.float() - exp).abs().max().item() < 1.5e-2
    assert not torch.equal(out, ref_out)


def test_missing_kernel_is_loud():
    q, k, v = _mk(1, 2, 2, 1, 256, 32, torch.bfloat16)  # head_dim 32: no sm_100a kernel
    with pytest.raises(RuntimeError):
        ta.tree_attention(q, k, v)


def test_pdl_back_to_back_steps_match_and_stay_ordered():
    """Programmatic dependent launch: consecutive steps overlap prologue/prefetch with the previous step's drain;
    q produced by a preceding kernel must still be honoured (griddepcontrol.wait before q is read)."""
    q, k, v = _mk(1, 16, 16, 1, 20000, 128, torch.bfloat16, seed=9)
    outs = []
    for i in range(20):
        qi = q * (1.0 + 0.05 * i)            # produced by the kernel right before the decode launch
        o, _ = L.decode_attention(qi, k, v, 0.088, impl="simt", pdl=2)
        outs.append((qi, o))
    torch.cuda.synchronize()
    for qi, o in outs[::5]:
        exp, _ = ref.attention_partial_ref(qi, k, v, 0.088, block=16384)
        assert (o.float() - exp).abs().max().item() < 1.5e-2


@pytest.mark.parametrize("host_io", ["zero_copy", "copy"])
def test_session_end_to_end_step_single_graph(host_io):
    """TreeDecodeSession.step, both host-I/O variants: "copy" = pinned-host query -> [H2D | fused attention | D2H] replayed
    as one CUDA graph -> pinned host result; "zero_copy" = the kernel itself loads q from / stores the result to pinned
    mapped host memory (one launch).  A new query every step, results equal the oracle."""
    from tree_attention_b200.models.decoder import TreeDecodeSession

    g = torch.Generator(device="cuda").manual_seed(5)
    k = torch.randn(1, 4, 3000, 128, device="cuda", generator=g).bfloat16()
    v = torch.randn(1, 4, 3000, 128, device="cuda", generator=g).bfloat16()
    sess = TreeDecodeSession([(k, v), (v, k)], softmax_scale=0.09, q_shape=(1, 8, 1, 128), host_io=host_io)
    oh = torch.empty(1, 8, 1, 128, dtype=torch.bfloat16).pin_memory()
    for it in range(4):
        q = torch.randn(1, 8, 1, 128, generator=torch.Generator().manual_seed(it)).bfloat16()
        qh = q.pin_memory()
        layer = it % 2
        got = sess.step(qh, oh, layer).clone()
        kk, vv = (k, v) if layer == 0 else (v, k)
        exp, _ = ref.attention_partial_ref(q.cuda(), kk, vv, 0.09)
        assert (got.float().cuda() - exp).abs().max().item() < 2e-2
    if host_io == "copy":
        assert len(sess.e2e_graphs) == 2 and not sess._steps_zc
    else:
        assert len(sess._steps_zc) == 2 and not sess.e2e_graphs


# ------------------------------------------------------------------------------------------------------------------
# device-resident fill level (kv_len): ADVICE r1 -- a preallocated, partially filled shard must only attend over the rows
# that were written; the kernels read the level on the device, so a captured graph follows a growing cache
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("impl", ["simt", "tc", "swap"])
@pytest.mark.parametrize("n_valid", [0, 1, 127, 128, 129, 1000, 2999, 3000])
def test_kv_len_masks_unwritten_rows(impl, n_valid):
    hq, hkv = (8, 8) if impl == "simt" else (8, 2)
    q, k, v = _mk(2, hq, hkv, 1, 3000, 128, torch.bfloat16, seed=11)
    # the unwritten tail: NaN for the CUDA-core kernel (it never touches those rows), huge-but-finite for the tensor-core
    # kernels (they multiply the tail of the last tile by exact zeros -- documented requirement: finite)
    junk = float("nan") if impl == "simt" else 3.0e4
    k[:, :, n_valid:] = junk
    v[:, :, n_valid:] = junk
    kv_len = torch.tensor([n_valid], dtype=torch.int32, device="cuda")
    out, lse = L.decode_attention(q, k, v, 0.088, False, 0, 0, impl=impl, kv_len=kv_len)
    torch.cuda.synchronize()
    if n_valid == 0:
        assert torch.all(out == 0) and torch.all(torch.isinf(lse) & (lse < 0))
        return
    o_ref, l_ref = ref.attention_partial_ref(q, k[:, :, :n_valid], v[:, :, :n_valid], 0.088, False, 0, 0, torch.float32)
    assert torch.isfinite(out).all()
    assert (out.float() - o_ref).abs().max().item() < 1.5e-2
    assert (lse - l_ref).abs().max().item() < 2e-3


@pytest.mark.parametrize("impl", ["simt", "swap"])
def test_kv_len_graph_follows_growing_cache(impl):
    """One captured graph, the fill level bumped on the device between replays."""
    hq, hkv = (4, 4) if impl == "simt" else (8, 2)
    q, k, v = _mk(1, hq, hkv, 1, 4096, 128, torch.bfloat16, seed=12)
    k[:, :, 100:] = 0
    v[:, :, 100:] = 0
    kv_len = torch.tensor([100], dtype=torch.int32, device="cuda")
    out = torch.empty_like(q)
    lse = torch.empty(1, hq, 1, device="cuda", dtype=torch.float32)
    L.decode_attention(q, k, v, 0.088, out=out, lse=lse, impl=impl, kv_len=kv_len)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g):
            L.decode_attention(q, k, v, 0.088, out=out, lse=lse, impl=impl, kv_len=kv_len)
    gen = torch.Generator(device="cuda").manual_seed(3)
    for n in (100, 101, 128, 129, 700, 4096):
        lo = int(kv_len.item())
        if n > lo:
            k[:, :, lo:n] = torch.randn(1, hkv, n - lo, 128, device="cuda", generator=gen).bfloat16()
            v[:, :, lo:n] = torch.randn(1, hkv, n - lo, 128, device="cuda", generator=gen).bfloat16()
        kv_len.fill_(n)
        g.replay()
        torch.cuda.synchronize()
        exp, _ = ref.attention_partial_ref(q, k[:, :, :n], v[:, :, :n], 0.088)
        assert (out.float() - exp).abs().max().item() < 1.5e-2, n


def test_session_fill_levels_append_and_prepared_launch():
    """TreeDecodeSession with fill levels on the native fast path (_C.DecodeStep): garbage past the fill level is ignored,
    append_kv raises the level on the device, graph replay (step) and prepared PDL launches (step_device) both follow."""
    from tree_attention_b200.models.decoder import TreeDecodeSession

    g = torch.Generator(device="cuda").manual_seed(21)
    cap, used = 2048, 1500
    k = torch.randn(1, 4, cap, 128, device="cuda", generator=g).bfloat16()
    v = torch.randn(1, 4, cap, 128, device="cuda", generator=g).bfloat16()
    k[:, :, used:] = float("nan")
    v[:, :, used:] = float("nan")
    sess = TreeDecodeSession([(k, v)], softmax_scale=0.09, q_shape=(1, 4, 1, 128), kv_lens=[used], pdl=True)
    q = torch.randn(1, 4, 1, 128, device="cuda", generator=g).bfloat16()
    oh = torch.empty(1, 4, 1, 128, dtype=torch.bfloat16).pin_memory()
    n = used
    for it in range(6):
        out_dev = sess.step_device(q, 0).clone()                   # prepared launch (PDL)
        out_e2e = sess.step(q.cpu().pin_memory(), oh, 0).clone()   # [H2D | kernel | D2H] graph
        exp, _ = ref.attention_partial_ref(q, k[:, :, :n], v[:, :, :n], 0.09)
        assert (out_dev.float() - exp).abs().max().item() < 1.5e-2, it
        assert (out_e2e.float().cuda() - exp).abs().max().item() < 1.5e-2, it
        k_new = torch.randn(1, 4, 3, 128, device="cuda", generator=g).bfloat16()
        v_new = torch.randn(1, 4, 3, 128, device="cuda", generator=g).bfloat16()
        sess.append_kv(0, k_new, v_new)
        n += 3
        assert sess.kv_len_host[0] == n
    assert sess._steps and sess._steps[0].impl == "simt"


def test_workspaces_are_per_stream():
    """Two streams decoding concurrently must not share partials / tickets."""
    q, k, v = _mk(1, 8, 8, 1, 20000, 128, torch.bfloat16, seed=13)
    q2, k2, v2 = _mk(1, 8, 8, 1, 20000, 128, torch.bfloat16, seed=14)
    exp1, _ = ref.attention_partial_ref(q, k, v, 0.088, block=16384)
    exp2, _ = ref.attention_partial_ref(q2, k2, v2, 0.088, block=16384)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    outs1, outs2 = [], []
    for _ in range(20):
        with torch.cuda.stream(s1):
            outs1.append(L.decode_attention(q, k, v, 0.088)[0])
        with torch.cuda.stream(s2):
            outs2.append(L.decode_attention(q2, k2, v2, 0.088)[0])
    torch.cuda.synchronize()
    for o in outs1:
        assert (o.float() - exp1).abs().max().item() < 1.5e-2
    for o in outs2:
        assert (o.float() - exp2).abs().max().item() < 1.5e-2
