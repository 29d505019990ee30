"""tcgen05 / TMEM / TMA flash-attention forward (``csrc/attn_fwd_sm100.cu``) -- Python wrapper.

Covers Sq >= 1 with any GQA ratio, causal masks with global offsets, head_dim 64/128, bf16/fp16.
``attention_fwd`` returns the shard-local partial ``(o, lse)``; ``attention_fwd_fused`` additionally runs
the cross-GPU combine over symmetric memory without NCCL.  Replaces the reference's local attention
(``/root/reference/model.py:74-80``) and, in fused mode, its three all-reduces (``model.py:105-116``).
"""
from __future__ import annotations

import os
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
    comm=None,
    variant: int = 0,
    comm_mode: int = 1,
    kv_seg: Optional[Tuple[int, int]] = None,
) -> Tuple[torch.Tensor, torch.Tensor]:
    """Shard-local partial ``(o, lse)``; with ``comm`` (a ``_C.Comm`` of the ``fwd`` family) the SAME launch also performs
    the cross-GPU combine: ``comm_mode=1`` -> the global result replicated on every rank, ``comm_mode=2`` -> the global
    result sharded over Sq (this rank's ``sq_out`` rows, see ``sharded_rows``).
    ``kv_seg=(seg_len, seg_gap)``: a two-segment shard (zigzag sharding): local rows ``>= seg_len`` sit ``seg_gap``
    positions further on in the global sequence (``seg_len % 128 == 0``, ``seg_gap >= 0``)."""
    C = _build.load()
    q = q if q.stride(-1) == 1 else q.contiguous()
    k = k if k.stride(-1) == 1 else k.contiguous()
    v = v if v.stride(-1) == 1 else v.contiguous()
    b, hq, sq, d = q.shape
    rows = sq
    if comm is not None and comm_mode == 2:
        rows = sharded_rows(sq, comm.world)
    if out is None:
        out = torch.empty((b, hq, rows, d), dtype=q.dtype, device=q.device)
    lse = torch.empty((b, hq, rows), dtype=torch.float32, device=q.device)
    C.attn_fwd(q, k, v, out, lse, float(softmax_scale), bool(causal), int(q_pos0), int(kv_pos0), comm,
               int(os.environ.get("TREE_ATTN_FWD_VARIANT", variant)), int(comm_mode),
               int(kv_seg[0]) if kv_seg else 0, int(kv_seg[1]) if kv_seg else 0)
    return out, lse


def sharded_rows(sq: int, world: int) -> int:
    """Rows of the Sq-sharded output per rank: whole 128-row query tiles, contiguous blocks in rank order."""
    tiles = (sq + 127) // 128
    return (tiles + world - 1) // world * 128


def bwd_eligible(q: torch.Tensor, k: torch.Tensor) -> bool:
    try:
        has = hasattr(_build.load(), "attn_bwd")
    except Exception:
        has = False
    return has and q.is_cuda and q.dtype in (torch.bfloat16, torch.float16) and q.shape[-1] in (64, 128)


def attention_bwd(q, k, v, o, lse, do, softmax_scale, causal=False, q_pos0=0, kv_pos0=0):
    """tcgen05 backward over THIS rank's KV shard with the global ``o``/``lse``.
    Returns ``(dq_partial fp32, dk, dv)``; ``dq_partial`` must be summed over ranks."""
    C = _build.load()
    q = q if q.stride(-1) == 1 else q.contiguous()
    k = k if k.stride(-1) == 1 else k.contiguous()
    v = v if v.stride(-1) == 1 else v.contiguous()
    o = o if o.stride(-1) == 1 else o.contiguous()
    do = do if do.stride(-1) == 1 else do.contiguous()
    b, hq, sq, d = q.shape
    hkv, s = k.shape[1], k.shape[2]
    sq_pad = (sq + 63) // 64 * 64
    dq = torch.empty((b, hq, sq, d), dtype=torch.float32, device=q.device)
    dk = torch.empty((b, hkv, s, d), dtype=q.dtype, device=q.device)
    dv = torch.empty((b, hkv, s, d), dtype=q.dtype, device=q.device)
    delta = torch.empty((b, hq, sq_pad), dtype=torch.float32, device=q.device)
    lse2 = torch.empty((b, hq, sq_pad), dtype=torch.float32, device=q.device)
    C.attn_bwd(q, k, v, o, do, lse.contiguous(), dq, dk, dv, delta, lse2, float(softmax_scale), bool(causal),
               int(q_pos0), int(kv_pos0))
    return dq, dk, dv


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
    output: str = "replicated",
    kv_seg: Optional[Tuple[int, int]] = None,
) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """ONE launch per rank: tcgen05 attention over the local KV shard; every 128-row query tile has an owner rank, the
    partial tiles are pushed to their owner over NVLink from the epilogue (reduce-scatter), merge CTAs of the same launch
    combine them and -- ``output="replicated"`` -- push the final tiles to every peer (all-gather).  No NCCL.

    ``output="sharded"``: returns this rank's rows of the global result, ``(B, Hq, sharded_rows(Sq, W), D)`` = rows
    ``[rank * n, (rank + 1) * n)`` (the last rank's block may run past Sq; those rows are undefined).  NVLink traffic per
    rank is ``(W-1)/W |O|`` (sharded) or twice that (replicated); the symmetric buffer is ``2 |O|`` either way.

    Replicated mode processes very long query blocks in chunks so that the symmetric buffer stays under
    ``TREE_ATTN_FWD_SYMM_CAP_GB`` (default 8)."""
    import os

    import torch.distributed as dist

    from ..parallel import symm

    C = _build.load()
    world = dist.get_world_size(group)
    b, hq, sq, d = q.shape
    cap = int(float(os.environ.get("TREE_ATTN_FWD_SYMM_CAP_GB", "8")) * (1 << 30))
    if output == "sharded":
        data, flags = C.attn_fwd_comm_bytes(b, hq, sq, d, world, 2)
        reg = symm.get_region("fwd_rs", int(data), int(flags), group, layout=(b, hq, sq, d))
        o_c, l_c = attention_fwd(q, k, v, softmax_scale, causal, q_pos0, kv_pos0, comm=reg.comm, comm_mode=2, kv_seg=kv_seg)
        return o_c, (l_c if return_lse else None)
    if output != "replicated":
        raise ValueError("output must be 'replicated' or 'sharded'")
    per_row = 2 * b * hq * (d * 2 + 4)   # partial slots on the owners + final slots: ~2 |O| bytes per query row
    chunk = max(128 * world, min(sq, (cap // per_row) // (128 * world) * (128 * world)))
    data, flags = C.attn_fwd_comm_bytes(b, hq, min(chunk, sq), d, world, 1)
    reg = symm.get_region("fwd", int(data), int(flags), group, layout=(b, hq, min(chunk, sq), d))
    out = torch.empty((b, hq, sq, d), dtype=q.dtype, device=q.device)
    lse = torch.empty((b, hq, sq), dtype=torch.float32, device=q.device)
    if chunk >= sq:
        o_c, l_c = attention_fwd(q, k, v, softmax_scale, causal, q_pos0, kv_pos0, out=out, comm=reg.comm, kv_seg=kv_seg)
        return o_c, l_c
    for s0 in range(0, sq, chunk):
        s1 = min(sq, s0 + chunk)
        reg = symm.get_region("fwd", int(data), int(flags), group, layout=(b, hq, s1 - s0, d))  # fences a ragged tail
        o_c, l_c = attention_fwd(q[:, :, s0:s1], k, v, softmax_scale, causal, q_pos0 + s0, kv_pos0,
                                 out=out[:, :, s0:s1], comm=reg.comm, kv_seg=kv_seg)
        lse[:, :, s0:s1] = l_c
    return out, lse
