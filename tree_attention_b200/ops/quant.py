"""Block-scaled fp8 (OCP MX style): e4m3 elements + one UE8M0 power-of-two scale per 32 elements of the
innermost dimension.  ``csrc/quant.cu`` holds the CUDA kernels; the PyTorch implementation below is the
oracle and the CPU path.  Used for the mxfp8 KV cache of the streaming decode kernel (half the HBM bytes of
bf16: decode is bandwidth bound, so this is up to 2x on the headline op).  The reference is fp16 only
(``/root/reference/model.py:51-53``)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple

import torch

from .. import _build

BLOCK = 32
E4M3_MAX = 448.0


def quantize_mxfp8_ref(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """PyTorch oracle: returns (uint8 e4m3 bytes, uint8 UE8M0 scales)."""
    assert x.shape[-1] % BLOCK == 0
    xf = x.float().reshape(*x.shape[:-1], x.shape[-1] // BLOCK, BLOCK)
    amax = xf.abs().amax(dim=-1)
    e = torch.ceil(torch.log2(torch.clamp(amax, min=1e-38) / E4M3_MAX))
    e = torch.where(amax > 0, e, torch.zeros_like(e)).clamp(-127, 127)
    q = (xf * torch.exp2(-e)[..., None]).clamp(-E4M3_MAX, E4M3_MAX).to(torch.float8_e4m3fn)
    return q.view(torch.uint8).reshape(x.shape), (e + 127).to(torch.uint8)


def dequantize_mxfp8_ref(q: torch.Tensor, scales: torch.Tensor) -> torch.Tensor:
    qf = q.view(torch.float8_e4m3fn).float().reshape(*q.shape[:-1], q.shape[-1] // BLOCK, BLOCK)
    s = torch.exp2(scales.float() - 127.0)
    return (qf * s[..., None]).reshape(q.shape)


def quantize_mxfp8(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    if x.is_cuda:
        return tuple(_build.load().quant_mxfp8(x.contiguous()))
    return quantize_mxfp8_ref(x)


def dequantize_mxfp8(q: torch.Tensor, scales: torch.Tensor) -> torch.Tensor:
    if q.is_cuda:
        return _build.load().dequant_mxfp8(q.contiguous(), scales.contiguous())
    return dequantize_mxfp8_ref(q, scales)


@dataclass
class MXFP8Tensor:
    """A block-scaled fp8 tensor: ``data`` uint8 (..., D) e4m3 bytes, ``scales`` uint8 (..., D/32) UE8M0."""

    data: torch.Tensor
    scales: torch.Tensor

    @classmethod
    def from_float(cls, x: torch.Tensor) -> "MXFP8Tensor":
        return cls(*quantize_mxfp8(x))

    @property
    def shape(self):
        return self.data.shape

    @property
    def is_cuda(self):
        return self.data.is_cuda

    @property
    def device(self):
        return self.data.device

    def dequantize(self, dtype: torch.dtype = torch.float32) -> torch.Tensor:
        return dequantize_mxfp8(self.data, self.scales).to(dtype)

    def write_rows(self, position: int, x: torch.Tensor) -> None:
        """KV append: quantise the new rows ``x`` (B, H, n, D) into ``[position, position + n)`` along the sequence
        (blocks run along the channels, so every row carries its own scales)."""
        q, sc = quantize_mxfp8(x.contiguous())
        n = x.shape[2]
        self.data[:, :, position:position + n].copy_(q, non_blocking=True)
        self.scales[:, :, position:position + n].copy_(sc, non_blocking=True)


SEQ_TILE = 128


@dataclass
class MXFP8SeqTensor:
    """Block-scaled fp8 with the 32-element blocks running along the SEQUENCE: ``data`` uint8 (B, H, S, D) e4m3,
    ``scales`` uint8 (B, H, ceil(S/128), D, 4) UE8M0 -- byte k of entry (tile, d) scales keys
    ``[128 tile + 32 k, 128 tile + 32 k + 32)`` of channel d.

    This is the V-cache format of the block-scaled tensor-core decode (``tcgen05.mma.kind::mxf8f6f4.block_scale`` in
    ``csrc/decode_swap_sm100.cu``): ``O^T += V^T P^T`` contracts over the keys, and MX scale factors apply per 32
    elements of the contraction, so V's blocks must run along the keys (K keeps the standard layout: ``S^T = K Q^T``
    contracts over the channels).  The four scale bytes of a (128-key tile, channel) pair form one 32-bit word, which is
    exactly one scale-factor entry of the MMA's A operand."""

    data: torch.Tensor
    scales: torch.Tensor

    @classmethod
    def from_float(cls, x: torch.Tensor) -> "MXFP8SeqTensor":
        if x.is_cuda and x.dtype in (torch.bfloat16, torch.float16, torch.float32):
            return cls(*_build.load().quant_mxfp8_seq(x.contiguous()))   # csrc/quant.cu
        return cls.from_float_ref(x)

    @classmethod
    def from_float_ref(cls, x: torch.Tensor) -> "MXFP8SeqTensor":
        """PyTorch oracle / CPU path of the quantiser."""
        b, h, s, d = x.shape
        t = (s + SEQ_TILE - 1) // SEQ_TILE
        xf = x.float()
        if t * SEQ_TILE != s:
            xf = torch.nn.functional.pad(xf, (0, 0, 0, t * SEQ_TILE - s))
        xb = xf.reshape(b, h, t, SEQ_TILE // BLOCK, BLOCK, d)
        amax = xb.abs().amax(dim=4)                                    # (B, H, T, 4, D)
        e = torch.ceil(torch.log2(torch.clamp(amax, min=1e-38) / E4M3_MAX))
        e = torch.where(amax > 0, e, torch.zeros_like(e)).clamp(-127, 127)
        q = (xb * torch.exp2(-e)[:, :, :, :, None, :]).clamp(-E4M3_MAX, E4M3_MAX).to(torch.float8_e4m3fn)
        data = q.view(torch.uint8).reshape(b, h, t * SEQ_TILE, d)[:, :, :s].contiguous()
        scales = (e + 127).to(torch.uint8).permute(0, 1, 2, 4, 3).contiguous()   # (B, H, T, D, 4)
        return cls(data, scales)

    @property
    def shape(self):
        return self.data.shape

    @property
    def is_cuda(self):
        return self.data.is_cuda

    @property
    def device(self):
        return self.data.device

    def write_rows(self, position: int, x: torch.Tensor) -> None:
        """KV append: rows ``[position, position + n)``.  A 32-key block shares one scale per channel, so the blocks the
        new rows touch are de-quantised, updated and re-quantised (at most 32 + n rows of work per call)."""
        n = x.shape[2]
        b, h, s, d = self.data.shape
        if self.data.is_cuda and x.dtype in (torch.bfloat16, torch.float16, torch.float32) and self.data.is_contiguous() \
                and self.scales.is_contiguous():
            _build.load().mxfp8_seq_append(self.data, self.scales, x, int(position))   # csrc/quant.cu: one small launch
            return
        self.write_rows_ref(position, x)

    def write_rows_ref(self, position: int, x: torch.Tensor) -> None:
        """PyTorch oracle / CPU path of ``write_rows``."""
        n = x.shape[2]
        b, h, s, d = self.data.shape
        lo = (position // BLOCK) * BLOCK
        hi = min(s, ((position + n + BLOCK - 1) // BLOCK) * BLOCK)
        blk = self.dequantize_rows(lo, hi)
        blk[:, :, position - lo:position - lo + n] = x.float()
        pad = (-(hi - lo)) % BLOCK
        if pad:
            blk = torch.nn.functional.pad(blk, (0, 0, 0, pad))
        xb = blk.reshape(b, h, -1, BLOCK, d)
        amax = xb.abs().amax(dim=3)                                    # (B, H, nblk, D)
        e = torch.ceil(torch.log2(torch.clamp(amax, min=1e-38) / E4M3_MAX))
        e = torch.where(amax > 0, e, torch.zeros_like(e)).clamp(-127, 127)
        q = (xb * torch.exp2(-e)[:, :, :, None, :]).clamp(-E4M3_MAX, E4M3_MAX).to(torch.float8_e4m3fn)
        self.data[:, :, lo:hi].copy_(q.view(torch.uint8).reshape(b, h, -1, d)[:, :, :hi - lo], non_blocking=True)
        sc = (e + 127).to(torch.uint8)
        for bi in range(sc.shape[2]):                                   # scale byte (tile, channel, block-in-tile)
            g = lo // BLOCK + bi
            self.scales[:, :, g // (SEQ_TILE // BLOCK), :, g % (SEQ_TILE // BLOCK)].copy_(sc[:, :, bi], non_blocking=True)

    def dequantize_rows(self, lo: int, hi: int) -> torch.Tensor:
        """fp32 rows ``[lo, hi)`` (``lo`` a multiple of 32)."""
        b, h, s, d = self.data.shape
        out = self.data[:, :, lo:hi].view(torch.float8_e4m3fn).float()
        for g in range(lo // BLOCK, (hi + BLOCK - 1) // BLOCK):
            sc = torch.exp2(self.scales[:, :, g // (SEQ_TILE // BLOCK), :, g % (SEQ_TILE // BLOCK)].float() - 127.0)
            r0, r1 = g * BLOCK - lo, min(hi, (g + 1) * BLOCK) - lo
            out[:, :, r0:r1] *= sc[:, :, None, :]
        return out

    def dequantize(self, dtype: torch.dtype = torch.float32) -> torch.Tensor:
        b, h, s, d = self.data.shape
        t = self.scales.shape[2]
        sc = torch.exp2(self.scales.float() - 127.0).permute(0, 1, 2, 4, 3)        # (B, H, T, 4, D)
        sc = sc[:, :, :, :, None, :].expand(b, h, t, SEQ_TILE // BLOCK, BLOCK, d).reshape(b, h, t * SEQ_TILE, d)[:, :, :s]
        return (self.data.view(torch.float8_e4m3fn).float() * sc).to(dtype)


# ------------------------------------------------------------------------------------------------
# per-channel-scaled fp8 (the tensor-core decode format)
# ------------------------------------------------------------------------------------------------
@dataclass
class FP8ChannelTensor:
    """e4m3 tensor with one fp32 scale per (batch, head, channel): ``data`` uint8 (B, H, S, D), ``scales`` (B, H, D).

    This is the KV-cache format of the tcgen05 decode kernel (``csrc/decode_tc_sm100.cu``, ``kind::f8f6f4``): K's
    channel scales fold into the query before it is quantised and V's apply to the output columns, so the inner
    loop is two plain fp8 GEMMs.  ``headroom`` > 1 leaves room for tokens appended later."""

    data: torch.Tensor
    scales: torch.Tensor

    @classmethod
    def from_float(cls, x: torch.Tensor, headroom: float = 1.0) -> "FP8ChannelTensor":
        amax = x.float().abs().amax(dim=2)  # (B, H, D)
        scales = torch.where(amax > 0, amax * (headroom / E4M3_MAX), torch.ones_like(amax)).contiguous()
        q = (x.float() / scales[:, :, None, :]).clamp(-E4M3_MAX, E4M3_MAX).to(torch.float8_e4m3fn)
        return cls(q.view(torch.uint8).contiguous(), scales)

    @property
    def shape(self):
        return self.data.shape

    @property
    def is_cuda(self):
        return self.data.is_cuda

    @property
    def device(self):
        return self.data.device

    def dequantize(self, dtype: torch.dtype = torch.float32) -> torch.Tensor:
        return (self.data.view(torch.float8_e4m3fn).float() * self.scales[:, :, None, :]).to(dtype)

    def write_rows(self, position: int, x: torch.Tensor) -> None:
        """KV append with the cache's channel scales (values beyond the scale's range saturate: build the cache with
        ``headroom`` > 1 when later tokens may be larger)."""
        q = (x.float() / self.scales[:, :, None, :]).clamp(-E4M3_MAX, E4M3_MAX).to(torch.float8_e4m3fn)
        self.data[:, :, position:position + x.shape[2]].copy_(q.view(torch.uint8), non_blocking=True)
