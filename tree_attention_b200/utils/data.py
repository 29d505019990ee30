"""Synthetic Q/K/V generation.

Reference parity: ``make_data(shape=(B, nh, T, C), rank, device)`` (``/root/reference/model.py:37-56``).
Differences, each a fix of a documented defect:

* D1 -- the reference documents ``(B, nh, T, C)`` but builds BSHD tensors; here the canonical
  layout is BHSD (``layout="bhsd"``) and ``layout="bshd"`` returns BSHD-*strided views of the
  same logical BHSD tensor*, which every op in this package accepts.
* D4 -- the reference seeds Q per rank, so ranks attend with different queries; here Q comes
  from a seed shared by all ranks and only the KV shard is seeded per rank.
* Data is generated directly on ``device`` (the reference generates fp32 on the host, casts,
  then copies: 4.5 s of its 9.9 s CPU run).

``T`` is the per-rank shard length, exactly as in the reference (global context = W x T).
"""
from __future__ import annotations

from typing import Tuple

import torch

from .logging import logger


def make_data(
    shape: Tuple[int, int, int, int],
    rank: int,
    device,
    dtype: torch.dtype = torch.float16,
    layout: str = "bhsd",
    sq: int = 1,
    num_kv_heads: int | None = None,
    seed: int = 0,
    log: bool = True,
):
    """Generate ``(Q, K, V)`` for rank ``rank``.

    ``shape = (B, nh, T, C)``: batch, query heads, per-rank KV length, head dim.
    Returns ``Q: (B, nh, sq, C)``, ``K, V: (B, nkv, T, C)`` in BHSD (or BSHD-strided views).
    """
    b, nh, t, c = shape
    nkv = nh if num_kv_heads is None else int(num_kv_heads)
    assert nh % nkv == 0
    device = torch.device(device)
    gq = torch.Generator(device=device)
    gq.manual_seed(seed)  # shared by every rank
    gkv = torch.Generator(device=device)
    gkv.manual_seed(seed + 1 + int(rank))  # per-rank shard

    def _randn(*sz, gen):
        # bf16/fp16 generation directly on device keeps 1M-token shards off the host.
        return torch.randn(*sz, device=device, dtype=torch.float32 if device.type == "cpu" else dtype,
                           generator=gen).to(dtype)

    if layout == "bhsd":
        q = _randn(b, nh, sq, c, gen=gq)
        k = _randn(b, nkv, t, c, gen=gkv)
        v = _randn(b, nkv, t, c, gen=gkv)
    elif layout == "bshd":
        q = _randn(b, sq, nh, c, gen=gq).transpose(1, 2)
        k = _randn(b, t, nkv, c, gen=gkv).transpose(1, 2)
        v = _randn(b, t, nkv, c, gen=gkv).transpose(1, 2)
    else:
        raise ValueError(f"unknown layout {layout!r}")
    if log:
        logger.info(f"Generated data on rank {rank} with shape {shape}.")
    return q, k, v
