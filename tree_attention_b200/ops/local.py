"""Local (single-shard) attention partial: ``(o, lse)`` of ``q`` against this rank's K/V.

Reference parity: ``flash_res_lse`` (``/root/reference/model.py:60-83``) -- same name, signature and
``(res, lse)`` contract, with ``lse`` the logsumexp of the scaled logits (fix of D2) and a true
``-inf`` causal mask (fix of D6).  On CUDA the work is done by the hand-written sm_100a kernels:

* ``decode``  -- ``csrc/decode_simt.cu``   (Sq x GQA-group small: HBM-bound streaming, split-KV)
* ``fwd``     -- ``csrc/attn_fwd_sm100.cu`` (tcgen05/TMEM/TMA flash forward)

and on CPU by the PyTorch oracle in ``ops/reference.py``.  There is no silent CUDA->PyTorch fallback:
if the extension is missing on a GPU machine, importing it raises.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from .. import _build
from . import reference as ref

_WS: Dict[tuple, Dict[str, torch.Tensor]] = {}

DECODE_MAX_ROWS = 128  # Sq * (Hq / Hkv): 1 row -> CUDA-core streaming kernel, 2..128 rows -> tcgen05 packed-tile kernel


def _as_bhsd(x: torch.Tensor) -> torch.Tensor:
    """Accept any strides as long as head_dim is contiguous (BHSD or BSHD-strided views)."""
    if x.stride(-1) != 1:
        x = x.contiguous()
    return x


def new_workspace(device: torch.device, nfloats: int, nticket: int) -> Dict[str, torch.Tensor]:
    """A private split-merge workspace (partials + tickets).  Objects that bake raw pointers into CUDA graphs or
    prepared launches (``TreeDecodeSession``) own one of these instead of borrowing the shared cache below."""
    # `part` holds tagged 8-byte words {fp32, launch tag} and `tickets` the 64-bit arrival counter the tags are derived
    # from (csrc/decode_comm.cuh): both start at zero, a tag is never zero, so stale memory can never pass for a word
    return {
        "part": torch.zeros(max(nfloats, 2), dtype=torch.float32, device=device),
        "tickets": torch.zeros(max(nticket, 64), dtype=torch.int32, device=device),
    }


def _workspace(device: torch.device, tag: str, nfloats: int, nticket: int) -> Dict[str, torch.Tensor]:
    """Shared workspace cache for eager calls, keyed by (device, STREAM, kernel family): two streams decoding
    concurrently never share partials / tickets, and growth is stream-ordered (the old buffers go back to the caching
    allocator, which only re-uses them behind the work already queued on that stream)."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream, tag)
    ws = _WS.get(key)
    if ws is None or ws["part"].numel() < nfloats or ws["tickets"].numel() < nticket:
        ws = new_workspace(device, nfloats, nticket)
        _WS[key] = ws
    return ws


def _kv_len_arg(kv_len, like: torch.Tensor):
    """``kv_len`` for the kernels: None, or a 1-element int32 tensor on the cache's device (python ints are wrapped)."""
    if kv_len is None:
        return None
    if isinstance(kv_len, torch.Tensor):
        if kv_len.dtype != torch.int32 or kv_len.device != like.device:
            kv_len = kv_len.to(device=like.device, dtype=torch.int32)
        return kv_len.reshape(1)
    return torch.tensor([int(kv_len)], dtype=torch.int32, device=like.device)


def decode_eligible(q: torch.Tensor, k: torch.Tensor) -> bool:
    if not q.is_cuda or q.dtype not in (torch.bfloat16, torch.float16):
        return False
    d = q.shape[-1]
    g = q.shape[1] // k.shape[1]
    return d in (64, 128) and q.shape[2] * g <= DECODE_MAX_ROWS


def decode_attention(
    q: torch.Tensor,
    k: torch.Tensor,
    v: torch.Tensor,
    softmax_scale: float,
    causal: bool = False,
    q_pos0: int = 0,
    kv_pos0: int = 0,
    comm=None,
    out: Optional[torch.Tensor] = None,
    lse: Optional[torch.Tensor] = None,
    return_lse: bool = True,
    impl: str = "auto",
    pdl: int = 0,
    kv_len=None,
) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """Launch the fused streaming decode kernel.  With ``comm`` (a ``_C.Comm``) the kernel also performs
    the cross-GPU tree combine and ``out``/``lse`` are the GLOBAL results, identical on every rank.

    ``kv_len``: number of VALID rows of this shard (python int or 1-element int32 device tensor; default: all
    ``k.shape[2]`` rows).  The kernels read it on the device at run time, so a CUDA graph captured once follows a
    growing KV cache; rows past ``kv_len`` never enter the softmax.  For the tensor-core kernels (``tc`` / ``swap``)
    the rows between ``kv_len`` and the end of its 128-row tile must hold finite values (zero-initialised caches do).

    ``impl``: ``"simt"`` (CUDA-core math; HBM-bound for one query row per KV head), ``"tc"`` (tcgen05: the
    (Hq/Hkv) x Sq rows of a KV head packed into one MMA tile; stays HBM-bound for GQA / multi-token decode),
    ``"swap"`` (tcgen05 swap-AB: keys on the TMEM lanes, <= 16 packed rows, head_dim 128; the fastest for small
    GQA groups), ``"auto"`` = simt for a single row, swap for 2..16 rows at head_dim 128, tc otherwise.
    ``pdl``: programmatic dependent launch for back-to-back decode steps (simt kernel): 1 = the next launch may
    start its prologue while this one drains; 2 = additionally prefetch K/V tiles before waiting on the previous
    kernel -- only valid when the KV cache was not written by the immediately preceding kernel of the stream."""
    import os

    C = _build.load()
    q, k, v = _as_bhsd(q), _as_bhsd(k), _as_bhsd(v)
    b, hq, sq, d = q.shape
    hkv, s = k.shape[1], k.shape[2]
    rows_total = (hq // hkv) * sq
    impl = os.environ.get("TREE_ATTN_DECODE_IMPL", impl)
    if impl == "auto":
        impl = "simt" if rows_total == 1 else ("swap" if (rows_total <= 16 and d == 128) else "tc")
    if impl in ("tc", "swap"):
        if q.stride(2) % 8 != 0 and sq > 1:
            q = q.contiguous()
        grid, max_parts, rows, part_floats, _ = C.decode_tc_plan(b, hq, hkv, sq, s, d)
        ws = _workspace(q.device, "decode_tc", part_floats, b * hkv + 2)
        if out is None:
            out = torch.empty((b, hq, sq, d), dtype=q.dtype, device=q.device)
        if lse is None and return_lse:
            lse = torch.empty((b, hq, sq), dtype=torch.float32, device=q.device)
        C.decode_tc_fwd(q, k, v, out, lse, ws["part"], ws["tickets"], comm, float(softmax_scale), bool(causal),
                        int(q_pos0), int(kv_pos0), impl == "swap", _kv_len_arg(kv_len, k))
        return out, lse
    grid, max_parts, rows, part_floats, _, _ = C.decode_plan(b, hq, hkv, sq, s, d)
    ws = _workspace(q.device, "decode", part_floats, b * hkv + 2)
    if out is None:
        out = torch.empty((b, hq, sq, d), dtype=q.dtype, device=q.device)
    if lse is None and return_lse:
        lse = torch.empty((b, hq, sq), dtype=torch.float32, device=q.device)
    C.decode_fwd(q, k, v, out, lse, ws["part"], ws["tickets"], comm, float(softmax_scale), bool(causal),
                 int(q_pos0), int(kv_pos0), int(os.environ.get("TREE_ATTN_PDL", pdl)), _kv_len_arg(kv_len, k))
    return out, lse


def decode_impl_for(q_shape, k_shape, impl: str = "auto") -> str:
    """Which decode kernel ``decode_attention`` picks for these shapes (bf16 / fp16 KV)."""
    import os

    b, hq, sq, d = q_shape
    rows_total = (hq // k_shape[1]) * sq
    impl = os.environ.get("TREE_ATTN_DECODE_IMPL", impl)
    if impl == "auto":
        impl = "simt" if rows_total == 1 else ("swap" if (rows_total <= 16 and d == 128) else "tc")
    return impl


def decode_workspace_sizes(q_shape, k_shape, impl: str) -> Tuple[int, int]:
    """(part floats, ticket ints) of the split-merge workspace for this problem."""
    C = _build.load()
    b, hq, sq, d = q_shape
    hkv, s = k_shape[1], k_shape[2]
    if impl in ("tc", "swap"):
        part_floats = C.decode_tc_plan(b, hq, hkv, sq, s, d)[3]
    else:
        part_floats = C.decode_plan(b, hq, hkv, sq, s, d)[3]
    return int(part_floats), b * hkv + 2


def decode_attention_mxfp8(
    q: torch.Tensor,
    k,
    v,
    softmax_scale: float,
    causal: bool = False,
    q_pos0: int = 0,
    kv_pos0: int = 0,
    comm=None,
    return_lse: bool = True,
    kv_len=None,
) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """Streaming decode over a block-scaled fp8 KV shard (``k``, ``v``: ``ops.quant.MXFP8Tensor``)."""
    C = _build.load()
    q = _as_bhsd(q)
    b, hq, sq, d = q.shape
    hkv, s = k.shape[1], k.shape[2]
    grid, max_parts, rows, part_floats, _, _ = C.decode_plan(b, hq, hkv, sq, s, d)
    ws = _workspace(q.device, "decode", part_floats, b * hkv + 2)
    out = torch.empty((b, hq, sq, d), dtype=q.dtype, device=q.device)
    lse = torch.empty((b, hq, sq), dtype=torch.float32, device=q.device) if return_lse else None
    C.decode_fwd_mx(q, k.data, v.data, k.scales, v.scales, out, lse, ws["part"], ws["tickets"], comm,
                    float(softmax_scale), bool(causal), int(q_pos0), int(kv_pos0), _kv_len_arg(kv_len, k.data))
    return out, lse


def decode_attention_fp8(
    q: torch.Tensor,
    k,
    v,
    softmax_scale: float,
    causal: bool = False,
    q_pos0: int = 0,
    kv_pos0: int = 0,
    comm=None,
    return_lse: bool = True,
    impl: str = "auto",
    kv_len=None,
) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """tcgen05 (kind::f8f6f4) decode over a per-channel-scaled e4m3 KV shard (``ops.quant.FP8ChannelTensor``).
    ``impl``: ``"swap"`` (keys on the TMEM lanes, <= 16 packed query rows) or ``"tc"`` (query rows on the lanes)."""
    C = _build.load()
    q = _as_bhsd(q)
    b, hq, sq, d = q.shape
    hkv, s = k.shape[1], k.shape[2]
    if q.stride(2) % 8 != 0 and sq > 1:
        q = q.contiguous()
    grid, max_parts, rows, part_floats, _ = C.decode_tc_plan(b, hq, hkv, sq, s, d)
    ws = _workspace(q.device, "decode_tc", part_floats, b * hkv + 2)
    out = torch.empty((b, hq, sq, d), dtype=q.dtype, device=q.device)
    lse = torch.empty((b, hq, sq), dtype=torch.float32, device=q.device) if return_lse else None
    import os

    impl = os.environ.get("TREE_ATTN_DECODE_FP8_IMPL", impl)
    if impl == "auto":
        impl = "swap" if rows <= 16 else "tc"
    C.decode_tc_fwd8(q, k.data, v.data, k.scales, v.scales, out, lse, ws["part"], ws["tickets"], comm,
                     float(softmax_scale), bool(causal), int(q_pos0), int(kv_pos0), impl == "swap", _kv_len_arg(kv_len, k.data))
    return out, lse


def decode_attention_mx_tc(
    q: torch.Tensor,
    k,
    v,
    softmax_scale: float,
    causal: bool = False,
    q_pos0: int = 0,
    kv_pos0: int = 0,
    comm=None,
    return_lse: bool = True,
    kv_len=None,
) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """Block-scaled fp8 KV cache on the tensor cores: ``k`` is an ``MXFP8Tensor`` (scales per 32 channels), ``v`` an
    ``MXFP8SeqTensor`` (scales per 32 keys); both GEMMs of the swap-AB decode kernel are
    ``tcgen05.mma.kind::mxf8f6f4.block_scale`` with the scale factors staged in TMEM.  head_dim 128, at most 16 packed
    query rows per KV head; with ``comm`` the kernel also performs the cross-GPU tree combine."""
    C = _build.load()
    q = _as_bhsd(q)
    b, hq, sq, d = q.shape
    hkv, s = k.data.shape[1], k.data.shape[2]
    if q.stride(2) % 8 != 0 and sq > 1:
        q = q.contiguous()
    grid, max_parts, rows, part_floats, _ = C.decode_tc_plan(b, hq, hkv, sq, s, d)
    ws = _workspace(q.device, "decode_tc", part_floats, b * hkv + 2)
    out = torch.empty((b, hq, sq, d), dtype=q.dtype, device=q.device)
    lse = torch.empty((b, hq, sq), dtype=torch.float32, device=q.device) if return_lse else None
    C.decode_mx_tc_fwd(q, k.data, v.data, k.scales.contiguous(), v.scales.contiguous(), out, lse, ws["part"], ws["tickets"],
                       comm, float(softmax_scale), bool(causal), int(q_pos0), int(kv_pos0), _kv_len_arg(kv_len, k.data))
    return out, lse


def decode_comm_bytes(b: int, hq: int, hkv: int, sq: int, s: int, d: int, world: int) -> Tuple[int, int]:
    """(data_bytes, flag_bytes) the decode family needs in symmetric memory (covers both decode kernels)."""
    rows = max(4, (hq // hkv) * sq)
    data = 2 * world * b * hkv * rows * (d + 2) * 8   # {fp32 value, epoch tag} words
    flags = 4096
    return data, flags


def attention_partial(
    q: torch.Tensor,
    k: torch.Tensor,
    v: torch.Tensor,
    softmax_scale: Optional[float] = None,
    causal: bool = False,
    q_pos0: int = 0,
    kv_pos0: int = 0,
    impl: str = "auto",
    kv_len=None,
) -> Tuple[torch.Tensor, torch.Tensor]:
    """Shard-local attention: returns ``(o, lse)`` with ``o`` in q's dtype (fp32 on CPU) and ``lse`` fp32.
    ``kv_len`` (valid rows of the shard) is honoured on the device by the decode kernels; every other path slices."""
    scale = ref.default_scale(q.shape[-1]) if softmax_scale is None else float(softmax_scale)
    if q.is_cuda:
        if impl in ("auto", "decode") and decode_eligible(q, k):
            return decode_attention(q, k, v, scale, causal, q_pos0, kv_pos0, kv_len=kv_len)
    if kv_len is not None:
        n = max(0, min(int(kv_len), k.shape[2]))   # (a device scalar is read back here: non-decode paths are not graph paths)
        if n == 0:
            o = torch.zeros(q.shape, dtype=q.dtype if q.is_cuda else torch.float32, device=q.device)
            return o, torch.full(q.shape[:-1], float("-inf"), dtype=torch.float32, device=q.device)
        k, v = k[:, :, :n], v[:, :, :n]
    if q.is_cuda:
        if impl in ("auto", "fwd"):
            from . import flash

            if flash.fwd_eligible(q, k):
                return flash.attention_fwd(q, k, v, scale, causal, q_pos0, kv_pos0)
        if impl == "torch":
            o, l = ref.attention_partial_ref(q, k, v, scale, causal, q_pos0, kv_pos0, torch.float32, block=8192)
            return o.to(q.dtype), l
        raise RuntimeError(
            f"no sm_100a kernel covers q={tuple(q.shape)} k={tuple(k.shape)} dtype={q.dtype}; "
            "pass impl='torch' explicitly to use the (slow) PyTorch path"
        )
    o, l = ref.attention_partial_ref(q, k, v, scale, causal, q_pos0, kv_pos0, torch.float32,
                                     block=8192 if q.shape[2] * k.shape[2] > (1 << 26) else 0)
    return o, l


def flash_res_lse(
    q: torch.Tensor,
    k: torch.Tensor,
    v: torch.Tensor,
    softmax_scale: float = 1.0,
    is_causal: bool = False,
) -> Tuple[torch.Tensor, torch.Tensor]:
    """Drop-in for the reference's ``flash_res_lse`` (model.py:60): ``(res, lse)`` for BHSD inputs.

    ``softmax_scale`` keeps the reference default of 1.0 for signature parity (D7); ``lse`` has shape
    ``res.shape[:-1]`` and is the logsumexp of the scaled logits.  With ``is_causal`` the queries are
    taken to be the LAST ``Sq`` positions of the key sequence.
    """
    q_pos0 = k.shape[2] - q.shape[2]
    res, lse = attention_partial(q, k, v, softmax_scale, is_causal, q_pos0, 0)
    return res, lse
