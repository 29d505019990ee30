"""Block-scaled fp8 (MX) quantisation oracle + the CPU path of tree_attention with an mxfp8 KV cache."""
import torch

import tree_attention_b200 as ta
from tree_attention_b200.ops import quant
from tree_attention_b200.ops import reference as ref


def test_quant_roundtrip_error_bound_and_no_saturation():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(3, 5, 256, generator=g) * torch.logspace(-3, 3, 5)[None, :, None]
    q, s = quant.quantize_mxfp8_ref(x)
    assert q.dtype == torch.uint8 and s.dtype == torch.uint8 and s.shape == (3, 5, 8)
    y = quant.dequantize_mxfp8_ref(q, s)
    blk_amax = x.reshape(3, 5, 8, 32).abs().amax(-1, keepdim=True)
    err = (x - y).reshape(3, 5, 8, 32).abs()
    assert (err <= blk_amax * 2 ** -3).all()          # e4m3: 3 mantissa bits after a power-of-two block scale
    assert (y.abs().reshape(3, 5, 8, 32).amax(-1, keepdim=True) <= blk_amax * 1.07).all()  # never clipped
    z = quant.quantize_mxfp8_ref(torch.zeros(2, 32))
    assert (quant.dequantize_mxfp8_ref(*z) == 0).all()


def test_tree_attention_accepts_mxfp8_kv_on_cpu():
    g = torch.Generator().manual_seed(1)
    q = torch.randn(1, 4, 1, 128, generator=g)
    k = torch.randn(1, 2, 64, 128, generator=g)
    v = torch.randn(1, 2, 64, 128, generator=g)
    kq, vq = quant.MXFP8Tensor.from_float(k), quant.MXFP8Tensor.from_float(v)
    out = ta.tree_attention(q, kq, vq)
    exp, _ = ref.attention_partial_ref(q, kq.dequantize(), vq.dequantize())
    assert torch.allclose(out, exp, atol=1e-5)
    full, _ = ref.attention_partial_ref(q, k, v)
    assert (out - full).abs().max() < 0.1  # quantisation error only


def test_seq_blocked_mx_roundtrip_and_cpu_path():
    """MXFP8SeqTensor: 32-key blocks per channel, scale words grouped per 128-key tile (the tensor-core V layout)."""
    import tree_attention_b200 as ta
    from tree_attention_b200.ops import reference as ref

    g = torch.Generator().manual_seed(5)
    v = torch.randn(2, 3, 300, 128, generator=g) * torch.logspace(-2, 2, 300)[:, None]
    vq = quant.MXFP8SeqTensor.from_float(v)
    assert vq.data.shape == (2, 3, 300, 128) and vq.scales.shape == (2, 3, 3, 128, 4)
    back = vq.dequantize()
    # e4m3 has 3 mantissa bits: relative error per element <= 2^-4 of its block maximum
    blk_max = torch.nn.functional.pad(v, (0, 0, 0, 84)).reshape(2, 3, 12, 32, 128).abs().amax(3, keepdim=True)
    err = (torch.nn.functional.pad(back - v, (0, 0, 0, 84)).reshape(2, 3, 12, 32, 128).abs() / blk_max.clamp(min=1e-30)).max()
    assert err <= 2 ** -4 + 1e-6
    assert (vq.scales[:, :, 2, :, 2:] == 127).all()   # blocks past the end of the sequence are scale 1
    q = torch.randn(2, 6, 1, 128, generator=g)
    k = torch.randn(2, 3, 300, 128, generator=g)
    kq = quant.MXFP8Tensor.from_float(k)
    out = ta.tree_attention(q, kq, vq)
    exp, _ = ref.attention_partial_ref(q, kq.dequantize(), back)
    assert (out - exp).abs().max().item() < 1e-4


def test_quantised_cache_append_rows():
    """write_rows on the three fp8 cache formats == quantising the updated float cache (same blocks / scales)."""
    g = torch.Generator().manual_seed(9)
    base = torch.randn(1, 2, 200, 128, generator=g)
    new = torch.randn(1, 2, 3, 128, generator=g) * 4.0
    upd = base.clone()
    upd[:, :, 70:73] = new
    # K layout: per-row blocks
    kq = quant.MXFP8Tensor.from_float(base)
    kq.write_rows(70, new)
    ref_k = quant.MXFP8Tensor.from_float(upd)
    assert torch.equal(kq.data, ref_k.data) and torch.equal(kq.scales, ref_k.scales)
    # V layout: 32-key blocks are re-quantised
    vq = quant.MXFP8SeqTensor.from_float(base)
    vq.write_rows(70, new)
    ref_v = quant.MXFP8SeqTensor.from_float(upd)
    assert (vq.dequantize() - ref_v.dequantize()).abs().max().item() <= 2 ** -3 * upd.abs().max().item()
    rows = vq.dequantize()[:, :, 70:73]
    assert (rows - new).abs().max().item() <= 2 ** -4 * new.abs().max().item() * 2
    # a write that crosses a block and a tile boundary
    vq2 = quant.MXFP8SeqTensor.from_float(base)
    big = torch.randn(1, 2, 40, 128, generator=g)
    vq2.write_rows(110, big)
    upd2 = base.clone()
    upd2[:, :, 110:150] = big
    # (re-quantising already rounded values may pick a different but equally valid power-of-two scale)
    assert (vq2.dequantize() - upd2).abs().max().item() <= 2 ** -3 * upd2.abs().max().item()
    assert (vq2.dequantize()[:, :, :96] - quant.MXFP8SeqTensor.from_float(base).dequantize()[:, :, :96]).abs().max().item() == 0
    # per-channel: saturating write with the cache's scales
    cq = quant.FP8ChannelTensor.from_float(base, headroom=8.0)
    cq.write_rows(70, new)
    assert (cq.dequantize()[:, :, 70:73] - new).abs().max().item() <= 2 ** -4 * 8.0 * base.abs().max().item()


def test_quantiser_properties_hypothesis():
    """For both MX layouts: bounded relative error per block, no saturation, and re-quantising a de-quantised tensor is
    lossless in value (idempotence)."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=25, deadline=None)
    @given(s=st.integers(1, 300), log_scale=st.floats(-20, 20), seed=st.integers(0, 10_000))
    def prop(s, log_scale, seed):
        g = torch.Generator().manual_seed(seed)
        x = torch.randn(1, 2, s, 64, generator=g) * (2.0 ** log_scale)
        for cls in (quant.MXFP8Tensor, quant.MXFP8SeqTensor):
            t = cls.from_float(x)
            y = t.dequantize()
            assert torch.isfinite(y).all()
            assert (y - x).abs().max().item() <= 2 ** -4 * x.abs().max().item() + 1e-30
            assert (t.data.view(torch.float8_e4m3fn).float().abs() <= 448).all()
            y2 = cls.from_float(y).dequantize()
            assert torch.equal(y2, y)

    prop()
