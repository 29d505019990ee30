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
