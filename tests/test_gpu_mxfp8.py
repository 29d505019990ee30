"""mxfp8 quant kernels and the block-scaled-fp8 KV decode kernel vs their PyTorch oracles."""
import pytest
import torch

pytestmark = pytest.mark.gpu

import tree_attention_b200 as ta
from tree_attention_b200.ops import local as L
from tree_attention_b200.ops import quant
from tree_attention_b200.ops import reference as ref


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
def test_quant_kernel_matches_reference(dtype):
    g = torch.Generator(device="cuda").manual_seed(0)
    x = (torch.randn(2, 3, 257, 128, device="cuda", generator=g) * 4).to(dtype)
    q, s = quant.quantize_mxfp8(x)
    q_ref, s_ref = quant.quantize_mxfp8_ref(x)
    assert torch.equal(s, s_ref)
    y, y_ref = quant.dequantize_mxfp8(q, s), quant.dequantize_mxfp8_ref(q_ref, s_ref)
    # identical up to round-to-nearest ties in the e4m3 conversion
    assert (y - y_ref).abs().max().item() <= (x.float().abs().max().item() * 2 ** -3)
    assert ((q != q_ref).float().mean().item()) < 1e-3
    assert torch.equal(quant.dequantize_mxfp8(q_ref, s_ref), y_ref)


CASES = [
    (1, 32, 32, 1, 4096, False),
    (2, 8, 8, 1, 300, False),
    (1, 32, 8, 1, 8192, False),     # GQA -> 4 rows
    (1, 8, 2, 2, 2000, True),       # 8 rows -> two passes, causal
    (1, 4, 4, 1, 70000, False),
]


@pytest.mark.parametrize("case", CASES, ids=[str(i) for i in range(len(CASES))])
def test_mxfp8_decode_matches_dequantised_oracle(case):
    b, hq, hkv, sq, s, causal = case
    g = torch.Generator(device="cuda").manual_seed(3)
    q = torch.randn(b, hq, sq, 128, device="cuda", generator=g).bfloat16()
    k = torch.randn(b, hkv, s, 128, device="cuda", generator=g).bfloat16()
    v = torch.randn(b, hkv, s, 128, device="cuda", generator=g).bfloat16()
    kq, vq = quant.MXFP8Tensor.from_float(k), quant.MXFP8Tensor.from_float(v)
    scale = 128 ** -0.5
    out, lse = L.decode_attention_mxfp8(q, kq, vq, scale, causal, s - sq, 0)
    torch.cuda.synchronize()
    o_ref, l_ref = ref.attention_partial_ref(q, kq.dequantize(), vq.dequantize(), scale, causal, s - sq, 0, torch.float32, block=16384)
    assert (out.float() - o_ref).abs().max().item() < 2e-2
    assert (lse - l_ref).abs().max().item() < 1e-2
    # and the quantisation error vs the bf16 cache stays small
    o_full, _ = ref.attention_partial_ref(q, k, v, scale, causal, s - sq, 0, torch.float32, block=16384)
    assert (out.float() - o_full).abs().max().item() < 0.1


def test_public_api_with_mxfp8_cache():
    g = torch.Generator(device="cuda").manual_seed(5)
    q = torch.randn(1, 16, 1, 128, device="cuda", generator=g).bfloat16()
    k = torch.randn(1, 16, 5000, 128, device="cuda", generator=g).bfloat16()
    v = torch.randn(1, 16, 5000, 128, device="cuda", generator=g).bfloat16()
    kq, vq = quant.MXFP8Tensor.from_float(k), quant.MXFP8Tensor.from_float(v)
    out = ta.tree_attention(q, kq, vq)
    exp, _ = ref.attention_partial_ref(q, kq.dequantize(), vq.dequantize())
    assert (out.float() - exp).abs().max().item() < 2e-2


FP8_CASES = [
    (1, 32, 32, 1, 4096, False),
    (2, 8, 8, 1, 300, False),
    (1, 32, 8, 1, 8192, False),     # GQA -> 4 packed rows
    (1, 8, 2, 4, 2000, True),       # 16 packed rows, causal
    (1, 4, 4, 1, 70000, False),
]


@pytest.mark.parametrize("impl", ["tc", "swap"])
@pytest.mark.parametrize("case", FP8_CASES, ids=[str(i) for i in range(len(FP8_CASES))])
def test_fp8_channel_tcgen05_decode_matches_dequantised_oracle(case, impl):
    """kind::f8f6f4 decode: per-channel-scaled e4m3 K/V, q quantised per row and P per element inside the kernel."""
    b, hq, hkv, sq, s, causal = case
    g = torch.Generator(device="cuda").manual_seed(11)
    q = torch.randn(b, hq, sq, 128, device="cuda", generator=g).bfloat16()
    k = torch.randn(b, hkv, s, 128, device="cuda", generator=g).bfloat16()
    v = torch.randn(b, hkv, s, 128, device="cuda", generator=g).bfloat16()
    kq, vq = quant.FP8ChannelTensor.from_float(k), quant.FP8ChannelTensor.from_float(v)
    scale = 128 ** -0.5
    out, lse = L.decode_attention_fp8(q, kq, vq, scale, causal, s - sq, 0, impl=impl)
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    o_ref, l_ref = ref.attention_partial_ref(q, kq.dequantize(), vq.dequantize(), scale, causal, s - sq, 0, torch.float32, block=16384)
    # q and P are additionally rounded to e4m3 inside the kernel: tolerance is fp8-level
    assert (out.float() - o_ref).abs().max().item() < 6e-2
    assert (lse - l_ref).abs().max().item() < 6e-2
    o_full, _ = ref.attention_partial_ref(q, k, v, scale, causal, s - sq, 0, torch.float32, block=16384)
    assert (out.float() - o_full).abs().max().item() < 0.12


def test_public_api_with_fp8_channel_cache():
    g = torch.Generator(device="cuda").manual_seed(12)
    q = torch.randn(1, 32, 1, 128, device="cuda", generator=g).bfloat16()
    k = torch.randn(1, 8, 6000, 128, device="cuda", generator=g).bfloat16()
    v = torch.randn(1, 8, 6000, 128, device="cuda", generator=g).bfloat16()
    kq, vq = quant.FP8ChannelTensor.from_float(k), quant.FP8ChannelTensor.from_float(v)
    out = ta.tree_attention(q, kq, vq)
    exp, _ = ref.attention_partial_ref(q, kq.dequantize(), vq.dequantize())
    assert (out.float() - exp).abs().max().item() < 6e-2


MX_TC_CASES = [
    (1, 32, 32, 1, 4096, False),
    (2, 8, 8, 1, 300, False),       # ragged last tile (scales padded to the 128-key tile)
    (1, 32, 8, 1, 8192, False),     # GQA -> 4 packed rows
    (1, 8, 2, 4, 2000, True),       # 16 packed rows, causal
    (1, 4, 4, 1, 70000, False),     # many tiles per CTA, split heads
    (1, 2, 1, 3, 129, True),        # one full + one 1-key tile
]


@pytest.mark.parametrize("case", MX_TC_CASES, ids=[str(i) for i in range(len(MX_TC_CASES))])
def test_mx_block_scaled_tcgen05_decode_matches_dequantised_oracle(case):
    """tcgen05.mma.kind::mxf8f6f4.block_scale decode: K blocks of 32 along the channels, V blocks of 32 along the keys,
    scale factors staged in TMEM; q MX-quantised and P rounded to e4m3 inside the kernel."""
    b, hq, hkv, sq, s, causal = case
    g = torch.Generator(device="cuda").manual_seed(21)
    q = torch.randn(b, hq, sq, 128, device="cuda", generator=g).bfloat16()
    # spread the magnitudes so that block scales differ along both the channels and the sequence
    k = (torch.randn(b, hkv, s, 128, device="cuda", generator=g) * torch.logspace(-1, 1, 128, device="cuda")).bfloat16()
    v = (torch.randn(b, hkv, s, 128, device="cuda", generator=g) * torch.logspace(-1.5, 1.5, s, device="cuda")[:, None]).bfloat16()
    kq, vq = quant.MXFP8Tensor.from_float(k), quant.MXFP8SeqTensor.from_float(v)
    scale = 128 ** -0.5 * 0.3
    out, lse = L.decode_attention_mx_tc(q, kq, vq, scale, causal, s - sq, 0)
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    vd = vq.dequantize()
    o_ref, l_ref = ref.attention_partial_ref(q, kq.dequantize(), vd, scale, causal, s - sq, 0, torch.float32, block=16384)
    # q (MX e4m3) and P (e4m3) carry 3 mantissa bits: ~3 % of the output range, independent of shape / masking
    tol = 6e-2 * max(1.0, o_ref.abs().max().item())
    assert (out.float() - o_ref).abs().max().item() < tol
    assert (lse - l_ref).abs().max().item() < 6e-2


def test_public_api_with_mx_block_scaled_cache():
    g = torch.Generator(device="cuda").manual_seed(22)
    q = torch.randn(1, 32, 1, 128, device="cuda", generator=g).bfloat16()
    k = torch.randn(1, 8, 6000, 128, device="cuda", generator=g).bfloat16()
    v = torch.randn(1, 8, 6000, 128, device="cuda", generator=g).bfloat16()
    kq, vq = quant.MXFP8Tensor.from_float(k), quant.MXFP8SeqTensor.from_float(v)
    out = ta.tree_attention(q, kq, vq)
    exp, _ = ref.attention_partial_ref(q, kq.dequantize(), vq.dequantize())
    assert (out.float() - exp).abs().max().item() < 6e-2


@pytest.mark.parametrize("fmt", ["mx_tc", "channel"])
def test_decode_session_with_quantised_cache_and_append(fmt):
    """TreeDecodeSession over an fp8 KV cache: graph-captured step, KV append (quantised on the way in), step again."""
    from tree_attention_b200.models.decoder import TreeDecodeSession

    g = torch.Generator(device="cuda").manual_seed(31)
    s, used = 4096, 4000
    q = torch.randn(1, 8, 1, 128, device="cuda", generator=g).bfloat16()
    k = torch.randn(1, 2, s, 128, device="cuda", generator=g).bfloat16()
    v = torch.randn(1, 2, s, 128, device="cuda", generator=g).bfloat16()
    k[:, :, used:] = 0
    v[:, :, used:] = 0
    if fmt == "mx_tc":
        kq, vq = quant.MXFP8Tensor.from_float(k), quant.MXFP8SeqTensor.from_float(v)
    else:
        kq, vq = quant.FP8ChannelTensor.from_float(k, headroom=4.0), quant.FP8ChannelTensor.from_float(v, headroom=4.0)
    scale = 128 ** -0.5
    sess = TreeDecodeSession([(kq, vq)], softmax_scale=scale, q_shape=(1, 8, 1, 128), dtype=torch.bfloat16)
    out0 = sess.step_device(q, 0).clone()
    exp0, _ = ref.attention_partial_ref(q, kq.dequantize(), vq.dequantize(), scale)
    assert (out0.float() - exp0).abs().max().item() < 6e-2
    k_new = torch.randn(1, 2, 2, 128, device="cuda", generator=g).bfloat16() * 2
    v_new = torch.randn(1, 2, 2, 128, device="cuda", generator=g).bfloat16() * 2
    sess.append_kv(0, k_new, v_new, used)
    out1 = sess.step_device(q, 0).clone()
    torch.cuda.synchronize()
    exp1, _ = ref.attention_partial_ref(q, kq.dequantize(), vq.dequantize(), scale)
    assert (out1.float() - exp1).abs().max().item() < 6e-2
    assert (vq.dequantize()[:, :, used:used + 2] - v_new.float()).abs().max().item() < 0.3
    assert (out1.float() - out0.float()).abs().max().item() > 1e-4   # the appended tokens are attended


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("s", [128, 300, 4097])
def test_seq_blocked_quantiser_kernel_matches_oracle(s, dtype):
    g = torch.Generator(device="cuda").manual_seed(s)
    x = (torch.randn(2, 3, s, 128, device="cuda", generator=g) * torch.logspace(-2, 2, s, device="cuda")[:, None]).to(dtype)
    a = quant.MXFP8SeqTensor.from_float(x)
    b = quant.MXFP8SeqTensor.from_float_ref(x)
    assert a.scales.shape == b.scales.shape and torch.equal(a.scales, b.scales)
    assert torch.equal(a.data, b.data)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("position,n", [(0, 1), (31, 1), (31, 2), (100, 70), (4090, 6), (4000, 96)])
def test_seq_blocked_append_kernel_matches_oracle(position, n, dtype):
    """Device-side KV append into the key-blocked MX cache (csrc/quant.cu mxfp8_seq_append_kernel) == the PyTorch oracle
    (de-quantise the touched 32-key blocks, insert, re-quantise), bit for bit."""
    g = torch.Generator(device="cuda").manual_seed(position * 7 + n)
    x = torch.randn(2, 3, 4096, 128, device="cuda", generator=g).to(dtype)
    a = quant.MXFP8SeqTensor.from_float(x)
    b = quant.MXFP8SeqTensor(a.data.clone(), a.scales.clone())
    new = (torch.randn(2, 3, n, 128, device="cuda", generator=g) * 4).to(dtype)
    a.write_rows(position, new)          # CUDA kernel
    b.write_rows_ref(position, new)      # oracle
    torch.cuda.synchronize()
    assert torch.equal(a.scales, b.scales)
    assert torch.equal(a.data, b.data)
    err = (a.dequantize()[:, :, position:position + n] - new.float()).abs()
    assert (err / new.float().abs().clamp(min=1.0)).max().item() < 0.07      # e4m3: 3 mantissa bits
