"""Pure-PyTorch oracle for the local attention partial and the (O, lse) combine.

This is the semantic specification every CUDA kernel in this package is tested
against, and the CPU / gloo execution path.

Reference parity (``/root/reference/model.py``):

* ``flash_res_lse`` (model.py:60-83) returns ``(res, lse)``.  The reference takes the
  logsumexp of the *probabilities* (model.py:80) and "masks" with ``torch.tril``
  (model.py:75-76); both are defects.  Here ``lse`` is the
  logsumexp of the scaled logits in fp32 and masked logits are ``-inf``.
* ``tree_decode`` (model.py:85-124) merges partials with (max, sum, sum, divide).
  ``merge_pair`` / ``merge_many`` are that formula written as an associative monoid on
  ``(o, lse)`` with identity ``(0, -inf)``.

Layout: canonical BHSD -- ``q: (B, Hq, Sq, D)``, ``k, v: (B, Hkv, Skv, D)`` with
``Hq % Hkv == 0`` (GQA/MQA).  All statistics are fp32.
"""
from __future__ import annotations

import math
from typing import Iterable, List, Optional, Sequence, Tuple

import torch

NEG_INF = float("-inf")


def default_scale(head_dim: int) -> float:
    return 1.0 / math.sqrt(head_dim)


def _expand_kv(x: torch.Tensor, n_rep: int) -> torch.Tensor:
    """(B, Hkv, S, D) -> (B, Hkv*n_rep, S, D) by repeating each kv head n_rep times."""
    if n_rep == 1:
        return x
    b, hkv, s, d = x.shape
    return x[:, :, None, :, :].expand(b, hkv, n_rep, s, d).reshape(b, hkv * n_rep, s, d)


def causal_mask(
    sq: int,
    skv: int,
    q_pos0: int,
    kv_pos0: int,
    device: torch.device,
) -> torch.Tensor:
    """Boolean (Sq, Skv) mask, True where key ``kv_pos0 + j`` is visible to query ``q_pos0 + i``."""
    qi = torch.arange(sq, device=device, dtype=torch.int64)[:, None] + int(q_pos0)
    kj = torch.arange(skv, device=device, dtype=torch.int64)[None, :] + int(kv_pos0)
    return kj <= qi


def attention_partial_ref(
    q: torch.Tensor,
    k: torch.Tensor,
    v: torch.Tensor,
    softmax_scale: Optional[float] = None,
    causal: bool = False,
    q_pos0: int = 0,
    kv_pos0: int = 0,
    compute_dtype: torch.dtype = torch.float32,
    block: int = 0,
) -> Tuple[torch.Tensor, torch.Tensor]:
    """Attention of ``q`` against ONE kv shard.

    Returns ``(o, lse)``: ``o`` is the shard-normalised output ``softmax(s) @ v`` in
    ``compute_dtype`` with shape ``(B, Hq, Sq, D)`` and ``lse`` is the row-wise logsumexp of
    the scaled logits, shape ``(B, Hq, Sq)``.  A row with every key masked yields the
    monoid identity ``(0, -inf)``.

    ``block > 0`` processes the keys in blocks with an online softmax so that very long
    shards never materialise the full ``(Sq, Skv)`` score matrix.
    """
    assert q.dim() == 4 and k.dim() == 4 and v.dim() == 4, "expected BHSD tensors"
    b, hq, sq, d = q.shape
    hkv, skv = k.shape[1], k.shape[2]
    assert hq % hkv == 0, f"Hq={hq} must be a multiple of Hkv={hkv}"
    scale = default_scale(d) if softmax_scale is None else float(softmax_scale)
    qf = q.to(compute_dtype)
    if block and skv > block:
        parts_o: List[torch.Tensor] = []
        parts_l: List[torch.Tensor] = []
        for s0 in range(0, skv, block):
            o_i, l_i = attention_partial_ref(
                q, k[:, :, s0 : s0 + block], v[:, :, s0 : s0 + block], scale, causal,
                q_pos0, kv_pos0 + s0, compute_dtype, 0,
            )
            parts_o.append(o_i)
            parts_l.append(l_i)
        return merge_many(parts_o, parts_l)
    kf = _expand_kv(k, hq // hkv).to(compute_dtype)
    vf = _expand_kv(v, hq // hkv).to(compute_dtype)
    s = torch.matmul(qf, kf.transpose(-2, -1)) * scale
    if causal:
        vis = causal_mask(sq, skv, q_pos0, kv_pos0, q.device)
        s = s.masked_fill(~vis, NEG_INF)
    lse = torch.logsumexp(s, dim=-1)
    dead = torch.isinf(lse) & (lse < 0)
    # exp(s - lse) with lse = -inf would be NaN; substitute 0 for dead rows.
    p = torch.exp(s - torch.where(dead, torch.zeros_like(lse), lse)[..., None])
    p = torch.where(dead[..., None], torch.zeros_like(p), p)
    o = torch.matmul(p, vf)
    return o, (lse if compute_dtype == torch.float64 else lse.to(torch.float32))


def merge_pair(
    o1: torch.Tensor, lse1: torch.Tensor, o2: torch.Tensor, lse2: torch.Tensor
) -> Tuple[torch.Tensor, torch.Tensor]:
    """Associative, commutative merge of two normalised partials.  Identity: ``(0, -inf)``."""
    m = torch.maximum(lse1, lse2)
    dead = torch.isinf(m) & (m < 0)
    m_safe = torch.where(dead, torch.zeros_like(m), m)
    w1 = torch.exp(lse1 - m_safe)
    w2 = torch.exp(lse2 - m_safe)
    den = w1 + w2
    den_safe = torch.where(dead, torch.ones_like(den), den)
    o = (o1 * w1[..., None].to(o1.dtype) + o2 * w2[..., None].to(o2.dtype)) / den_safe[..., None].to(o1.dtype)
    lse = torch.where(dead, torch.full_like(m, NEG_INF), m_safe + torch.log(den_safe))
    return o, lse


def merge_many(
    os: Sequence[torch.Tensor], lses: Sequence[torch.Tensor]
) -> Tuple[torch.Tensor, torch.Tensor]:
    """Flat (max, sum, sum, divide) merge -- the reference's formulation (model.py:108-124)."""
    assert len(os) == len(lses) and len(os) > 0
    if len(os) == 1:
        return os[0], lses[0]
    lse_all = torch.stack(list(lses), dim=0)
    o_all = torch.stack(list(os), dim=0)
    m = lse_all.max(dim=0).values
    dead = torch.isinf(m) & (m < 0)
    m_safe = torch.where(dead, torch.zeros_like(m), m)
    w = torch.exp(lse_all - m_safe[None])
    den = w.sum(dim=0)
    den_safe = torch.where(dead, torch.ones_like(den), den)
    num = (o_all * w[..., None].to(o_all.dtype)).sum(dim=0)
    o = num / den_safe[..., None].to(o_all.dtype)
    lse = torch.where(dead, torch.full_like(m, NEG_INF), m_safe + torch.log(den_safe))
    return o, lse


def merge_tree(
    os: Sequence[torch.Tensor], lses: Sequence[torch.Tensor]
) -> Tuple[torch.Tensor, torch.Tensor]:
    """Binary-tree pairwise merge in fixed rank order (depth ceil(log2 W))."""
    items = list(zip(os, lses))
    while len(items) > 1:
        nxt = []
        for i in range(0, len(items) - 1, 2):
            nxt.append(merge_pair(items[i][0], items[i][1], items[i + 1][0], items[i + 1][1]))
        if len(items) % 2:
            nxt.append(items[-1])
        items = nxt
    return items[0]


def attention_ref(
    q: torch.Tensor,
    k: torch.Tensor,
    v: torch.Tensor,
    softmax_scale: Optional[float] = None,
    causal: bool = False,
    q_pos0: Optional[int] = None,
    compute_dtype: torch.dtype = torch.float64,
) -> Tuple[torch.Tensor, torch.Tensor]:
    """Monolithic attention over the full (concatenated) KV -- the test oracle."""
    if q_pos0 is None:
        q_pos0 = k.shape[2] - q.shape[2]
    return attention_partial_ref(q, k, v, softmax_scale, causal, q_pos0, 0, compute_dtype)


def sharded_attention_ref(
    q: torch.Tensor,
    k_shards: Iterable[torch.Tensor],
    v_shards: Iterable[torch.Tensor],
    softmax_scale: Optional[float] = None,
    causal: bool = False,
    q_pos0: Optional[int] = None,
    schedule: str = "flat",
) -> Tuple[torch.Tensor, torch.Tensor]:
    """Single-process emulation of W ranks (a "fake group"): partial per shard + merge."""
    k_shards, v_shards = list(k_shards), list(v_shards)
    total = sum(int(ks.shape[2]) for ks in k_shards)
    if q_pos0 is None:
        q_pos0 = total - q.shape[2]
    os, lses, off = [], [], 0
    for ks, vs in zip(k_shards, v_shards):
        o, l = attention_partial_ref(q, ks, vs, softmax_scale, causal, q_pos0, off)
        os.append(o)
        lses.append(l)
        off += int(ks.shape[2])
    return (merge_tree if schedule == "tree" else merge_many)(os, lses)


def attention_bwd_ref(
    q: torch.Tensor,
    k: torch.Tensor,
    v: torch.Tensor,
    o: torch.Tensor,
    lse: torch.Tensor,
    do: torch.Tensor,
    softmax_scale: Optional[float] = None,
    causal: bool = False,
    q_pos0: int = 0,
    kv_pos0: int = 0,
) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Backward of one KV shard given the GLOBAL ``o`` and ``lse``.

    ``dK``/``dV`` of the shard are complete; the returned ``dQ`` is this shard's partial and
    must be summed over shards.  fp32 math.
    """
    b, hq, sq, d = q.shape
    hkv = k.shape[1]
    g = hq // hkv
    scale = default_scale(d) if softmax_scale is None else float(softmax_scale)
    qf, of, dof = q.float(), o.float(), do.float()
    kf = _expand_kv(k, g).float()
    vf = _expand_kv(v, g).float()
    s = torch.matmul(qf, kf.transpose(-2, -1)) * scale
    if causal:
        vis = causal_mask(sq, k.shape[2], q_pos0, kv_pos0, q.device)
        s = s.masked_fill(~vis, NEG_INF)
    dead = torch.isinf(lse) & (lse < 0)
    lse_safe = torch.where(dead, torch.zeros_like(lse), lse)
    p = torch.exp(s - lse_safe[..., None])
    p = torch.where(dead[..., None], torch.zeros_like(p), p)
    delta = (dof * of).sum(dim=-1)
    dv = torch.matmul(p.transpose(-2, -1), dof)
    dp = torch.matmul(dof, vf.transpose(-2, -1))
    ds = p * (dp - delta[..., None]) * scale
    dq = torch.matmul(ds, kf)
    dk = torch.matmul(ds.transpose(-2, -1), qf)
    if g > 1:
        skv = k.shape[2]
        dk = dk.reshape(b, hkv, g, skv, d).sum(dim=2)
        dv = dv.reshape(b, hkv, g, skv, d).sum(dim=2)
    return dq, dk, dv
