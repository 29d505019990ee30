"""tcgen05 layout conventions, pinned down on hardware with a one-CTA GEMM (csrc/umma_probe.cu)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from tree_attention_b200 import _build


@pytest.mark.parametrize("a_from_tmem", [False, True])
@pytest.mark.parametrize("b_mn_major", [False, True])
@pytest.mark.parametrize("n,k", [(128, 64), (128, 128), (64, 128), (256, 128), (128, 256), (16, 128), (32, 64)])
def test_umma_probe(n, k, b_mn_major, a_from_tmem):
    if b_mn_major and n % 64:
        pytest.skip("MN-major B needs N % 64 == 0")
    C = _build.load()
    g = torch.Generator(device="cuda").manual_seed(n * 1000 + k)
    a = torch.randn(128, k, device="cuda", generator=g).bfloat16()
    b = torch.randn((k, n) if b_mn_major else (n, k), device="cuda", generator=g).bfloat16()
    c = torch.zeros(128, n, device="cuda", dtype=torch.float32)
    C.umma_probe(a, b, c, b_mn_major, a_from_tmem)
    torch.cuda.synchronize()
    exp = a.float() @ (b.float() if b_mn_major else b.float().t())
    err = (c - exp).abs().max().item()
    assert err < 1e-2 * max(1.0, exp.abs().max().item() / 16), f"max err {err}"


@pytest.mark.parametrize("a_mn_major", [False, True])
@pytest.mark.parametrize("n", [16, 64, 128])
def test_umma_block_scaled_probe(n, a_mn_major):
    """tcgen05.mma.kind::mxf8f6f4.block_scale with UE8M0 scale factors staged in TMEM (32 lanes x 4 columns per 128
    rows, replicated per lane quarter, byte k of a word = K-block k) vs the de-quantised PyTorch product."""
    from tree_attention_b200.ops import quant

    C = _build.load()
    g = torch.Generator(device="cuda").manual_seed(n)
    a = torch.randn(128, 128, device="cuda", generator=g) * torch.logspace(-2, 2, 128, device="cuda")[:, None]
    b = torch.randn(n, 128, device="cuda", generator=g) * 3
    a8, sfa = quant.quantize_mxfp8(a.contiguous())
    b8, sfb = quant.quantize_mxfp8(b.contiguous())
    c = torch.zeros(128, n, device="cuda", dtype=torch.float32)
    # MN-major A: the operand is stored transposed ([K][M]) -- how a [key][channel] V tile is consumed along the keys
    C.umma_bs_probe(a8.t().contiguous() if a_mn_major else a8, b8, sfa.contiguous(), sfb.contiguous(), c, a_mn_major)
    torch.cuda.synchronize()
    exp = quant.dequantize_mxfp8(a8, sfa) @ quant.dequantize_mxfp8(b8, sfb).t()
    err = (c - exp).abs().max().item() / exp.abs().max().item()
    assert err < 1e-5, err


@pytest.mark.skipif(not __import__("os").environ.get("TREE_ATTN_EXPERIMENTAL"),
                    reason="cta_group::2 probe is compile-checked only so far (docs/NEXT.md); set TREE_ATTN_EXPERIMENTAL=1 to run it")
def test_umma_2cta_probe():
    """One tcgen05.mma.cta_group::2 GEMM (M = 256 over a CTA pair, each CTA holding half of B's rows)."""
    C = _build.load()
    g = torch.Generator(device="cuda").manual_seed(3)
    a = torch.randn(256, 64, device="cuda", generator=g).bfloat16()
    b = torch.randn(128, 64, device="cuda", generator=g).bfloat16()
    c = torch.zeros(256, 128, device="cuda", dtype=torch.float32)
    C.umma_2cta_probe(a, b, c)
    torch.cuda.synchronize()
    exp = a.float() @ b.float().t()
    assert (c - exp).abs().max().item() < 1e-2
