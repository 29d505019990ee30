"""Tree Attention: sequence-parallel attention over a KV sequence sharded across ranks.

Reference parity: ``tree_decode(q, k, v, rank, world_size, device)`` (``/root/reference/model.py:85-124``)
and the ``tree_attention()`` entry point named by BASELINE.json.  Q is replicated, every rank owns a
contiguous KV shard, each rank computes a local partial ``(o_r, lse_r)`` and the partials are merged
with the exact associative combine ``out = sum_r o_r e^{lse_r - m} / sum_r e^{lse_r - m}``.

Backends
--------
``fused``   one hand-written sm_100a kernel per rank: local attention + split merge + cross-GPU combine
            over symmetric memory (P2P stores, release/acquire epoch flags).  No NCCL on the path.
``symm``    local attention kernel, then the stand-alone ``combine_partials`` kernel over symmetric
            memory (one-shot or butterfly).  No NCCL on the path.
``nccl`` / ``gloo``  local partial + ``torch.distributed`` collectives: the reference's structure
            (``schedule="allreduce3"`` = MAX, SUM, SUM exactly as model.py:108-115, but with an
            ``|O| + 2``-scalar payload instead of ``3|O|``), or one packed ``allgather``, or a real
            pairwise ``butterfly`` tree, or a sequential ``ring`` chain (depth 2(W-1): the pattern the Tree Attention
            paper compares against).  This is the baseline and the CPU plumbing path.
``local``   world size 1.
"""
from __future__ import annotations

from typing import Optional, Tuple, Union

import torch
import torch.distributed as dist

from ..ops import local as local_ops
from ..ops import reference as ref
from . import symm

import contextlib
import os as _os


def _nvtx(name: str):
    """NVTX range around the op when TREE_ATTN_NVTX=1 (profilers only; nothing is logged in the hot path)."""
    if _os.environ.get("TREE_ATTN_NVTX") and torch.cuda.is_available():
        return torch.cuda.nvtx.range(name)
    return contextlib.nullcontext()


_BACKENDS = ("auto", "fused", "symm", "nccl", "gloo", "collective", "local")
_SCHEDULES = ("oneshot", "butterfly", "allreduce3", "allgather", "ring")


def _world(group) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def _resolve_backend(backend: str, q: torch.Tensor, world: int) -> str:
    if backend not in _BACKENDS:
        raise ValueError(f"backend must be one of {_BACKENDS}")
    if world == 1:
        return "local"
    if backend in ("nccl", "gloo"):
        return "collective"
    if backend == "auto":
        return "fused" if q.is_cuda else "collective"
    return backend


# ------------------------------------------------------------------------------------------------
# collective-based combines (baseline / CPU)
# ------------------------------------------------------------------------------------------------
def _combine_allreduce3(o: torch.Tensor, lse: torch.Tensor, group) -> Tuple[torch.Tensor, torch.Tensor]:
    """MAX, SUM, SUM -- the reference's schedule (model.py:108,114,115) with fp32 statistics and a
    per-row (not per-element) max/denominator payload (fixes D8/D9)."""
    o32 = o.float()
    m = lse.clone()
    dist.all_reduce(m, op=dist.ReduceOp.MAX, group=group)
    dead = torch.isinf(m) & (m < 0)
    m_safe = torch.where(dead, torch.zeros_like(m), m)
    w = torch.exp(lse - m_safe)
    num = o32 * w[..., None]
    den = w.clone()
    dist.all_reduce(num, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(den, op=dist.ReduceOp.SUM, group=group)
    den_safe = torch.where(dead, torch.ones_like(den), den)
    out = num / den_safe[..., None]
    lse_g = torch.where(dead, torch.full_like(m, float("-inf")), m_safe + torch.log(den_safe))
    return out, lse_g


def _combine_allgather(o: torch.Tensor, lse: torch.Tensor, group) -> Tuple[torch.Tensor, torch.Tensor]:
    """One packed all-gather of (o, lse), then a local merge in fixed rank order (deterministic)."""
    world = dist.get_world_size(group)
    packed = torch.cat([o.float(), lse[..., None]], dim=-1).contiguous()
    gathered = [torch.empty_like(packed) for _ in range(world)]
    dist.all_gather(gathered, packed, group=group)
    os_ = [g[..., :-1] for g in gathered]
    ls_ = [g[..., -1] for g in gathered]
    return ref.merge_many(os_, ls_)


def _combine_butterfly(o: torch.Tensor, lse: torch.Tensor, group) -> Tuple[torch.Tensor, torch.Tensor]:
    """log2(W) rounds of pairwise exchange; both partners apply merge(lower, higher) => identical bits."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if world & (world - 1):
        return _combine_allgather(o, lse, group)
    cur = torch.cat([o.float(), lse[..., None]], dim=-1).contiguous()
    k = 1
    while k < world:
        partner = rank ^ k
        other = torch.empty_like(cur)
        gp = dist.get_global_rank(group, partner) if group is not None else partner
        ops = [dist.P2POp(dist.isend, cur, gp, group), dist.P2POp(dist.irecv, other, gp, group)]
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        lo, hi = (cur, other) if rank < partner else (other, cur)
        o_m, l_m = ref.merge_pair(lo[..., :-1], lo[..., -1], hi[..., :-1], hi[..., -1])
        cur = torch.cat([o_m, l_m[..., None]], dim=-1).contiguous()
        k <<= 1
    return cur[..., :-1], cur[..., -1]


def _combine_ring(o, lse, group):
    """Sequential chain: rank 0 -> 1 -> ... -> W-1 accumulates the partials in rank order (W-1 dependent hops), then the
    result travels back along the chain.  This is the communication pattern a ring (blockwise) attention would use to move a
    decode query's state through the KV shards -- depth ``2 (W-1)`` against ``ceil(log2 W)`` for the butterfly and 1 for the
    fused one-shot: the comparison the Tree Attention paper draws.  Baseline only; never the product path."""
    rank, world = _world(group)
    g = (lambda r: dist.get_global_rank(group, r)) if group is not None else (lambda r: r)
    acc = torch.cat([o.float(), lse.float()[..., None]], dim=-1).contiguous()
    if rank > 0:
        prev = torch.empty_like(acc)
        dist.recv(prev, src=g(rank - 1), group=group)
        o_m, l_m = ref.merge_pair(prev[..., :-1], prev[..., -1], acc[..., :-1], acc[..., -1])   # earlier ranks first
        acc = torch.cat([o_m, l_m[..., None]], dim=-1).contiguous()
    if rank < world - 1:
        dist.send(acc, dst=g(rank + 1), group=group)
        dist.recv(acc, src=g(rank + 1), group=group)                        # the global result on its way back
    if rank > 0:
        dist.send(acc, dst=g(rank - 1), group=group)
    return acc[..., :-1], acc[..., -1]


def combine_partials(
    o: torch.Tensor,
    lse: torch.Tensor,
    group=None,
    backend: str = "auto",
    schedule: str = "oneshot",
    out_dtype: Optional[torch.dtype] = None,
) -> Tuple[torch.Tensor, torch.Tensor]:
    """Merge per-rank partials ``(o, lse)`` across ``group``; the result is replicated on every rank."""
    rank, world = _world(group)
    out_dtype = out_dtype or o.dtype
    if world == 1:
        return o.to(out_dtype), lse
    use_symm = o.is_cuda and backend in ("auto", "symm", "fused") and schedule != "ring"   # ring exists as a collective baseline only
    if use_symm:
        from .. import _build

        C = _build.load()
        mode = 1 if schedule == "butterfly" and (world & (world - 1)) == 0 else 0
        d = o.shape[-1]
        rows = o.numel() // d
        nchunks = (rows + 7) // 8
        import math

        rounds = max(1, int(math.ceil(math.log2(world))))
        slots = world if mode == 0 else rounds
        reg = symm.get_region("combine", 2 * slots * rows * (d + 4) * 4, 2 * slots * nchunks * 4, group,
                              layout=(mode, rows, d))
        o32 = o.float().contiguous()
        lse32 = lse.float().contiguous()
        out = torch.empty(o.shape, dtype=out_dtype if out_dtype in (torch.float32, torch.bfloat16, torch.float16)
                          else torch.float32, device=o.device)
        lse_out = torch.empty_like(lse32)
        C.combine(o32, lse32, out, lse_out, reg.comm, mode)
        return out.to(out_dtype), lse_out
    if schedule == "allreduce3":
        out, lse_g = _combine_allreduce3(o, lse, group)
    elif schedule == "ring":
        out, lse_g = _combine_ring(o, lse, group)
    elif schedule == "butterfly":
        out, lse_g = _combine_butterfly(o, lse, group)
    else:
        out, lse_g = _combine_allgather(o, lse, group)
    return out.to(out_dtype), lse_g


def allreduce_sum(x: torch.Tensor, group=None, backend: str = "auto") -> torch.Tensor:
    """Sum ``x`` (fp32) over the ranks of ``group``.  On CUDA this is ONE hand-written kernel over symmetric
    memory (pull reduce-scatter + push all-gather, ``csrc/reduce.cu``) -- the backward's dQ tree reduce; the
    collective path (NCCL / gloo ``all_reduce``) is the baseline and the CPU path.  Returns a new tensor."""
    rank, world = _world(group)
    if world == 1:
        return x
    if x.is_cuda and backend in ("auto", "symm", "fused") and x.dtype == torch.float32 and x.numel() % 4 == 0:
        import os

        from .. import _build

        C = _build.load()
        xc = x.contiguous()
        y = torch.empty_like(xc)
        flat_x, flat_y = xc.view(-1), y.view(-1)
        cap = int(float(os.environ.get("TREE_ATTN_REDUCE_SYMM_CAP_GB", "4")) * (1 << 30))
        chunk = max(4096, (cap // 16) // 4096 * 4096)  # floats per launch: 16 B of symmetric memory per float
        n = flat_x.numel()
        data, flags = C.symm_allreduce_sizes(min(n, chunk), world)
        for s0 in range(0, n, chunk):
            s1 = min(n, s0 + chunk)
            reg = symm.get_region("reduce", int(data), int(flags), group, layout=(s1 - s0,))
            C.symm_allreduce(flat_x[s0:s1], flat_y[s0:s1], reg.comm)
        return y
    y = x.clone()
    dist.all_reduce(y, op=dist.ReduceOp.SUM, group=group)
    return y


# ------------------------------------------------------------------------------------------------
# public API
# ------------------------------------------------------------------------------------------------
def tree_attention(
    q: torch.Tensor,
    k: torch.Tensor,
    v: torch.Tensor,
    *,
    group=None,
    causal: bool = False,
    softmax_scale: Optional[float] = None,
    kv_offset: Optional[int] = None,
    q_offset: Optional[int] = None,
    return_lse: bool = False,
    backend: str = "auto",
    schedule: str = "oneshot",
    layout: str = "bhsd",
    decode_pdl: int = 0,
    kv_len=None,
    output: str = "replicated",
    kv_layout: str = "contiguous",
) -> Union[torch.Tensor, Tuple[torch.Tensor, torch.Tensor]]:
    """Exact attention of replicated ``q`` over a KV sequence sharded across the ranks of ``group``.

    q: ``(B, Hq, Sq, D)``; k, v: ``(B, Hkv, S_local, D)`` -- this rank's contiguous shard (``layout="bshd"``
    takes ``(B, S, H, D)`` tensors instead).  ``softmax_scale=None`` means ``1/sqrt(D)``.

    Causal masking uses GLOBAL positions: key ``j`` of this shard sits at ``kv_offset + j`` (default
    ``rank * S_local``: equal contiguous shards) and query ``i`` at ``q_offset + i`` (default: the last
    ``Sq`` positions of the global sequence, i.e. decode / chunked-prefill convention).

    ``decode_pdl`` (single-row decode only): launch with programmatic dependent launch so that back-to-back
    decode steps overlap one step's drain (peer wait) with the next step's prologue; ``2`` also prefetches K/V
    before the dependency wait and requires that the KV cache was not written by the preceding kernel.

    ``kv_len``: number of VALID rows of this rank's shard (python int or 1-element int32 device tensor; default: the
    whole shard).  A preallocated, partially filled KV cache passes its fill level here: rows past it never enter
    the softmax.  The decode kernels read the value on the device, so a CUDA graph captured once follows a growing
    cache; ranks may have different (including zero) fill levels.

    ``output="sharded"``: return only this rank's block of query rows of the global result (and ``lse``):
    ``n = ceil(ceil(Sq / 128) / W) * 128`` rows ``[rank * n, (rank + 1) * n)`` as a ``(B, Hq, n, D)`` tensor (rows past
    ``Sq`` in the last block are undefined).  On the fused prefill path this is a reduce-scatter inside the attention
    kernel -- every rank receives ``(W-1)/W |O|`` bytes instead of sending ``(W-1) |O|`` -- other paths slice.

    ``kv_layout="zigzag"`` (causal prefill): this rank's shard is ``[chunk r | chunk 2W-1-r]`` of the sequence cut into
    ``2W`` equal chunks (``zigzag_shard``).  With contiguous shards a causal mask leaves the last rank almost idle while
    rank 0 does the full work (every rank waits for it in the combine); zigzag gives every rank ``(2W+1)/(4W)`` of that.
    The tcgen05 forward kernel takes the two-segment shard natively (one launch, fused combine included); other paths
    compute one partial per segment and merge.  Non-causal calls ignore the layout (positions do not matter).

    Returns the global attention output (replicated, bitwise identical across ranks for the fused and
    symm backends) and optionally the global ``lse``.
    """
    from ..ops.quant import FP8ChannelTensor, MXFP8Tensor

    if output not in ("replicated", "sharded"):
        raise ValueError("output must be 'replicated' or 'sharded'")
    if output == "sharded" and layout != "bhsd":
        raise ValueError("output='sharded' needs layout='bhsd'")

    quantised = isinstance(k, (MXFP8Tensor, FP8ChannelTensor))
    if layout == "bshd":
        if quantised:
            raise ValueError("quantised KV caches are stored (B, H, S, D): pass q as (B, H, Sq, D) with layout='bhsd'")
        q, k, v = q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)
    elif layout != "bhsd":
        raise ValueError("layout must be 'bhsd' or 'bshd'")
    rank, world = _world(group)
    scale = ref.default_scale(q.shape[-1]) if softmax_scale is None else float(softmax_scale)
    s_local = k.shape[2]

    if quantised and kv_layout == "zigzag" and causal and world > 1:
        raise ValueError("kv_layout='zigzag' is not supported for quantised KV caches")
    if quantised:  # fp8 KV cache (block-scaled MX or per-channel scaled)
        return _tree_attention_mxfp8(q, k, v, group, rank, world, scale, causal, kv_offset, q_offset, return_lse,
                                     backend, schedule, kv_len)
    if kv_layout not in ("contiguous", "zigzag"):
        raise ValueError("kv_layout must be 'contiguous' or 'zigzag'")
    if kv_layout == "zigzag" and causal and world > 1:
        if kv_offset is not None or kv_len is not None or quantised:
            raise ValueError("kv_layout='zigzag' derives the key positions itself: kv_offset / kv_len / quantised caches are not supported")
        o, lse = _tree_attention_zigzag(q, k, v, group, rank, world, scale, q_offset, return_lse, backend, schedule, output)
        if layout == "bshd":
            o = o.transpose(1, 2)
        return (o, lse) if return_lse else o
    kv_pos0 = rank * s_local if kv_offset is None else int(kv_offset)
    q_pos0 = (world * s_local - q.shape[2]) if q_offset is None else int(q_offset)
    be = _resolve_backend(backend, q, world)

    if be == "local":
        with _nvtx("tree_attention/local"):
            if decode_pdl and local_ops.decode_eligible(q, k):
                o, lse = local_ops.decode_attention(q, k, v, scale, causal, q_pos0, kv_pos0, pdl=decode_pdl, kv_len=kv_len)
            else:
                o, lse = local_ops.attention_partial(q, k, v, scale, causal, q_pos0, kv_pos0, kv_len=kv_len)
    elif be == "fused":
        if not q.is_cuda:
            raise RuntimeError("backend='fused' needs CUDA tensors")
        if local_ops.decode_eligible(q, k):
            b, hq, sq, d = q.shape
            data, flags = local_ops.decode_comm_bytes(b, hq, k.shape[1], sq, s_local, d, world)
            rows = (hq // k.shape[1]) * sq
            reg = symm.get_region("decode" if rows == 1 else "decode_tc", data, flags, group,
                                  layout=(b, hq, k.shape[1], sq, d))
            o, lse = local_ops.decode_attention(q, k, v, scale, causal, q_pos0, kv_pos0, comm=reg.comm,
                                                return_lse=return_lse, pdl=decode_pdl, kv_len=kv_len)
        else:
            from ..ops import flash

            if kv_len is not None:
                raise NotImplementedError("kv_len with the fused prefill kernel: slice the shard instead (every rank "
                                          "must still launch; an empty shard is not supported on this path)")
            o, lse = flash.attention_fwd_fused(q, k, v, scale, causal, q_pos0, kv_pos0, group=group,
                                               return_lse=return_lse, output=output)
            if output == "sharded":   # already this rank's rows (reduce-scatter inside the kernel)
                return (o, lse) if return_lse else o
    else:  # symm | collective
        with _nvtx("tree_attention/local_partial"):
            o_p, lse_p = local_ops.attention_partial(q, k, v, scale, causal, q_pos0, kv_pos0, kv_len=kv_len)
        sched = schedule
        if be == "collective" and schedule == "oneshot":
            sched = "allgather"
        with _nvtx(f"tree_attention/combine[{be}:{sched}]"):
            o, lse = combine_partials(o_p, lse_p, group, "symm" if be == "symm" else "collective", sched,
                                      out_dtype=o_p.dtype)
    if output == "sharded":
        o, lse = shard_rows(o, lse, rank, world)
    if layout == "bshd":
        o = o.transpose(1, 2)
    return (o, lse) if return_lse else o


def shard_rows(o: torch.Tensor, lse: Optional[torch.Tensor], rank: int, world: int):
    """This rank's block of query rows of a replicated ``(B, Hq, Sq, D)`` result, padded to whole 128-row tiles per rank
    (the layout the fused reduce-scatter path produces)."""
    sq = o.shape[2]
    n = ((sq + 127) // 128 + world - 1) // world * 128
    lo, hi = min(rank * n, sq), min((rank + 1) * n, sq)
    o_s = o.new_zeros((o.shape[0], o.shape[1], n, o.shape[3]))
    o_s[:, :, : hi - lo] = o[:, :, lo:hi]
    l_s = None
    if lse is not None:
        l_s = lse.new_full((lse.shape[0], lse.shape[1], n), float("-inf"))
        l_s[:, :, : hi - lo] = lse[:, :, lo:hi]
    return o_s, l_s


def zigzag_chunks(rank: int, world: int) -> Tuple[int, int]:
    """The two chunk indices (of ``2 * world`` equal chunks of the sequence) rank ``rank`` owns under zigzag sharding."""
    return rank, 2 * world - 1 - rank


def zigzag_shard(x: torch.Tensor, rank: int, world: int, dim: int = 2) -> torch.Tensor:
    """This rank's zigzag shard ``[chunk r | chunk 2W-1-r]`` of a full-sequence tensor (sequence along ``dim``, length a
    multiple of ``2 * world``)."""
    n = x.shape[dim]
    if n % (2 * world):
        raise ValueError(f"zigzag sharding needs a sequence length divisible by 2 * world ({n} % {2 * world} != 0)")
    c = n // (2 * world)
    a, b = zigzag_chunks(rank, world)
    return torch.cat([x.narrow(dim, a * c, c), x.narrow(dim, b * c, c)], dim)


def zigzag_unshard(shards, dim: int = 2) -> torch.Tensor:
    """Inverse of ``zigzag_shard``: the full sequence from the list of all ranks' shards (rank order)."""
    world = len(shards)
    c = shards[0].shape[dim] // 2
    chunks = [None] * (2 * world)
    for r, sh in enumerate(shards):
        a, b = zigzag_chunks(r, world)
        chunks[a], chunks[b] = sh.narrow(dim, 0, c), sh.narrow(dim, c, c)
    return torch.cat(chunks, dim)


def _tree_attention_zigzag(q, k, v, group, rank, world, scale, q_offset, return_lse, backend, schedule, output):
    """Causal attention over a zigzag-sharded KV sequence (``tree_attention(kv_layout="zigzag")``)."""
    s_local = k.shape[2]
    if s_local % 2:
        raise ValueError("kv_layout='zigzag' needs an even number of local KV rows (two equal chunks)")
    half = s_local // 2
    ca, cb = zigzag_chunks(rank, world)
    pos_a, pos_b = ca * half, cb * half
    gap = pos_b - pos_a - half            # >= 0: the second chunk always lies further on in the sequence
    q_pos0 = (world * s_local - q.shape[2]) if q_offset is None else int(q_offset)
    be = _resolve_backend(backend, q, world)
    native = False
    if q.is_cuda and half % 128 == 0 and not local_ops.decode_eligible(q, k):
        from ..ops import flash

        native = flash.fwd_eligible(q, k)
    if be == "fused" and native:
        with _nvtx("tree_attention/zigzag[fused]"):
            o, lse = flash.attention_fwd_fused(q, k, v, scale, True, q_pos0, pos_a, group=group, return_lse=return_lse,
                                               output=output, kv_seg=(half, gap))
        return o, lse
    with _nvtx("tree_attention/zigzag[local_partial]"):
        if native:   # one launch over both segments
            o_p, lse_p = flash.attention_fwd(q, k, v, scale, True, q_pos0, pos_a, kv_seg=(half, gap))
        else:
            parts = [local_ops.attention_partial(q, k[:, :, a:a + half], v[:, :, a:a + half], scale, True, q_pos0, pos)
                     for a, pos in ((0, pos_a), (half, pos_b))]
            o_p, lse_p = ref.merge_many([p[0].float() for p in parts], [p[1] for p in parts])
            o_p = o_p.to(parts[0][0].dtype)
    sched = "allgather" if (be == "collective" and schedule == "oneshot") else schedule
    with _nvtx("tree_attention/zigzag[combine]"):
        o, lse = combine_partials(o_p, lse_p, group, "collective" if be == "collective" else "symm", sched, out_dtype=o_p.dtype)
    if output == "sharded":
        o, lse = shard_rows(o, lse, rank, world)
    return o, lse


def _tree_attention_mxfp8(q, k, v, group, rank, world, scale, causal, kv_offset, q_offset, return_lse, backend,
                          schedule, kv_len=None):
    """mxfp8 KV: fused streaming decode when eligible (CUDA, head_dim 128, few query rows); otherwise dequantise."""
    s_local = k.shape[2]
    kv_pos0 = rank * s_local if kv_offset is None else int(kv_offset)
    q_pos0 = (world * s_local - q.shape[2]) if q_offset is None else int(q_offset)
    g = q.shape[1] // k.shape[1]
    from ..ops.quant import FP8ChannelTensor, MXFP8SeqTensor

    per_channel = isinstance(k, FP8ChannelTensor)
    mx_tc = isinstance(v, MXFP8SeqTensor)   # block-scaled tensor-core path: K blocks along channels, V blocks along keys
    max_rows = local_ops.DECODE_MAX_ROWS if per_channel else 16
    if (q.is_cuda and q.shape[-1] == 128 and q.shape[2] * g <= max_rows and backend in ("auto", "fused", "local")):
        comm = None
        if world > 1:
            b, hq, sq, d = q.shape
            data, flags = local_ops.decode_comm_bytes(b, hq, k.shape[1], sq, s_local, d, world)
            fam = "decode_tc" if (per_channel or mx_tc) else "decode"
            comm = symm.get_region(fam, data, flags, group, layout=("fp8", per_channel, mx_tc, b, hq, k.shape[1], sq, d)).comm
        fn = (local_ops.decode_attention_fp8 if per_channel else
              local_ops.decode_attention_mx_tc if mx_tc else local_ops.decode_attention_mxfp8)
        o, lse = fn(q, k, v, scale, causal, q_pos0, kv_pos0, comm=comm, return_lse=return_lse, kv_len=kv_len)
        return (o, lse) if return_lse else o
    dt = q.dtype if q.is_cuda else torch.float32
    return tree_attention(q, k.dequantize(dt), v.dequantize(dt), group=group, causal=causal, softmax_scale=scale,
                          kv_offset=kv_pos0, q_offset=q_pos0, return_lse=return_lse, backend=backend, schedule=schedule,
                          kv_len=kv_len)


def tree_decode(
    q: torch.Tensor,
    k: torch.Tensor,
    v: torch.Tensor,
    rank: int = 0,
    world_size: int = 1,
    device=None,
    *,
    softmax_scale: float = 1.0,
    backend: str = "auto",
    schedule: str = "oneshot",
    group=None,
) -> torch.Tensor:
    """Signature-compatible shim for the reference's ``tree_decode`` (model.py:85).

    ``rank``/``world_size``/``device`` are accepted for parity (the reference only uses them for log
    lines and the distributed-branch gate); the process group decides the real topology.  Like the
    reference it applies NO ``1/sqrt(d)`` scaling by default (model.py:60,100 -- D7) and is non-causal.
    Accepts BHSD ``q: (B, nh, 1, C)``, ``k, v: (B, nh, T, C)``.
    """
    _, world = _world(group)
    if world_size > 1 and world == 1:
        raise RuntimeError(
            f"tree_decode called with world_size={world_size} but no process group is initialised; "
            "call setup(rank, world_size) first"
        )
    return tree_attention(q, k, v, group=group, causal=False, softmax_scale=softmax_scale, backend=backend,
                          schedule=schedule)
