"""Autograd for tree attention (the reference, ``/root/reference/model.py``, has no backward at all -- SURVEY.md 5.7 / 7.4).

Forward leaves the GLOBAL ``o`` and ``lse`` on every rank.  Backward on rank r:

    delta = rowsum(dO * O)                                     (local)
    dK_r, dV_r = flash backward over the LOCAL kv shard using the GLOBAL lse   (no communication)
    dQ = sum_r dQ_r                                            (plain sum -> all-reduce / symmetric-memory reduce)

``dQ_r`` partials are fp32 and reduced in fp32.
"""
from __future__ import annotations

from typing import Optional

import torch

from ..parallel.tree import _world, tree_attention
from . import reference as ref


def _local_bwd(q, k, v, o, lse, do, scale, causal, q_pos0, kv_pos0):
    if q.is_cuda:
        from . import flash

        if hasattr(flash, "attention_bwd") and flash.bwd_eligible(q, k):
            return flash.attention_bwd(q, k, v, o, lse, do, scale, causal, q_pos0, kv_pos0)
    # blockwise PyTorch fallback (CPU path; also the oracle of the CUDA kernel)
    blk = 4096
    skv = k.shape[2]
    if skv <= blk:
        return ref.attention_bwd_ref(q, k, v, o, lse, do, scale, causal, q_pos0, kv_pos0)
    dq = torch.zeros(q.shape, dtype=torch.float32, device=q.device)
    dks, dvs = [], []
    for s0 in range(0, skv, blk):
        dq_i, dk_i, dv_i = ref.attention_bwd_ref(q, k[:, :, s0:s0 + blk], v[:, :, s0:s0 + blk], o, lse, do, scale,
                                                 causal, q_pos0, kv_pos0 + s0)
        dq += dq_i
        dks.append(dk_i)
        dvs.append(dv_i)
    return dq, torch.cat(dks, 2), torch.cat(dvs, 2)


class _TreeAttentionFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, causal, softmax_scale, group, kv_offset, q_offset, backend, schedule, kv_layout="contiguous"):
        rank, world = _world(group)
        scale = ref.default_scale(q.shape[-1]) if softmax_scale is None else float(softmax_scale)
        s_local = k.shape[2]
        q_pos0 = (world * s_local - q.shape[2]) if q_offset is None else int(q_offset)
        zigzag = kv_layout == "zigzag" and causal and world > 1
        if zigzag:
            # two equal chunks per rank: chunk r and chunk 2W-1-r of the sequence (parallel/tree.py zigzag_shard)
            from ..parallel.tree import zigzag_chunks

            half = s_local // 2
            segs = [(0, half, zigzag_chunks(rank, world)[0] * half), (half, half, zigzag_chunks(rank, world)[1] * half)]
            o, lse = tree_attention(q, k, v, group=group, causal=True, softmax_scale=scale, q_offset=q_pos0, return_lse=True,
                                    backend=backend, schedule=schedule, kv_layout="zigzag")
        else:
            kv_pos0 = rank * s_local if kv_offset is None else int(kv_offset)
            segs = [(0, s_local, kv_pos0)]
            o, lse = tree_attention(q, k, v, group=group, causal=causal, softmax_scale=scale, kv_offset=kv_pos0,
                                    q_offset=q_pos0, return_lse=True, backend=backend, schedule=schedule)
        ctx.save_for_backward(q, k, v, o, lse)
        ctx.meta = (scale, causal, q_pos0, segs, group, world)
        return o

    @staticmethod
    def backward(ctx, do):
        q, k, v, o, lse, = ctx.saved_tensors
        scale, causal, q_pos0, segs, group, world = ctx.meta
        do = do.contiguous()
        dq, dks, dvs = None, [], []
        for row0, n, pos in segs:   # one local backward per contiguous segment of the shard (two under zigzag sharding)
            dq_i, dk_i, dv_i = _local_bwd(q, k[:, :, row0:row0 + n], v[:, :, row0:row0 + n], o, lse, do, scale, causal, q_pos0, pos)
            dq = dq_i.float() if dq is None else dq + dq_i.float()
            dks.append(dk_i)
            dvs.append(dv_i)
        dk = dks[0] if len(dks) == 1 else torch.cat(dks, 2)
        dv = dvs[0] if len(dvs) == 1 else torch.cat(dvs, 2)
        dq = dq.contiguous()
        if world > 1:
            from ..parallel.tree import allreduce_sum

            dq = allreduce_sum(dq, group)  # symmetric-memory kernel on CUDA, all_reduce on CPU
        return dq.to(q.dtype), dk.to(k.dtype), dv.to(v.dtype), None, None, None, None, None, None, None, None


def tree_attention_func(
    q: torch.Tensor,
    k: torch.Tensor,
    v: torch.Tensor,
    causal: bool = False,
    softmax_scale: Optional[float] = None,
    group=None,
    kv_offset: Optional[int] = None,
    q_offset: Optional[int] = None,
    backend: str = "auto",
    schedule: str = "oneshot",
    layout: str = "bhsd",
    kv_layout: str = "contiguous",
) -> torch.Tensor:
    """Differentiable ``tree_attention`` (replicated q, sequence-sharded k/v; ``kv_layout="zigzag"``: balanced causal shards)."""
    if layout == "bshd":
        q, k, v = q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)
    if kv_layout == "zigzag" and kv_offset is not None:
        raise ValueError("kv_layout='zigzag' derives the key positions itself: do not pass kv_offset")
    o = _TreeAttentionFn.apply(q, k, v, causal, softmax_scale, group, kv_offset, q_offset, backend, schedule, kv_layout)
    return o.transpose(1, 2) if layout == "bshd" else o
