"""tcgen05 / TMEM / TMA flash-attention forward (``csrc/attn_fwd_sm100.cu``) -- Python wrapper.

Covers Sq >= 1 with any GQA ratio, causal masks with global offsets, head_dim 64/128, bf16/fp16.
``attention_fwd`` returns the shard-local partial ``(o, lse)``; ``attention_fwd_fused`` additionally runs
the cross-GPU combine over symmetric memory without NCCL.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from .. import _build


def _has_kernel() -> bool:
    try:
        return hasattr(_build.load(), "attn_fwd")
    except Exception:
        return False


def fwd_eligible(q: torch.Tensor, k: torch.Tensor) -> bool:
    return (
        q.is_cuda
        and q.dtype in (torch.bfloat16, torch.float16)
        and q.shape[-1] in (64, 128)
        and q.shape[1] % k.shape[1] == 0
        and _has_kernel()
    )


def attention_fwd(
    q: torch.Tensor,
    k: torch.Tensor,
    v: torch.Tensor,
    softmax_scale: float,
    causal: bool = False,
    q_pos0: int = 0,
    kv_pos0: int = 0,
    out: Optional[torch.Tensor] = None,
) -> Tuple[torch.Tensor, torch.Tensor]:
    C = _build.load()
    q = q if q.stride(-1) == 1 else q.contiguous()
    k = k if k.stride(-1) == 1 else k.contiguous()
    v = v if v.stride(-1) == 1 else v.contiguous()
    b, hq, sq, d = q.shape
    if out is None:
        out = torch.empty((b, hq, sq, d), dtype=q.dtype, device=q.device)
    lse = torch.empty((b, hq, sq), dtype=torch.float32, device=q.device)
    C.attn_fwd(q, k, v, out, lse, float(softmax_scale), bool(causal), int(q_pos0), int(kv_pos0))
    return out, lse


def attention_fwd_fused(
    q: torch.Tensor,
    k: torch.Tensor,
    v: torch.Tensor,
    softmax_scale: float,
    causal: bool = False,
    q_pos0: int = 0,
    kv_pos0: int = 0,
    group=None,
    return_lse: bool = True,
) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """Local tcgen05 forward + cross-GPU combine over symmetric memory (no NCCL)."""
    from ..parallel.tree import combine_partials

    o_p, lse_p = attention_fwd(q, k, v, softmax_scale, causal, q_pos0, kv_pos0)
    o, lse = combine_partials(o_p, lse_p, group, "symm", "oneshot", out_dtype=q.dtype)
    return o, lse
