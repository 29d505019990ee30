"""``torch.library`` registration of the shard-local attention op.

``torch.ops.tree_attention.partial(q, k, v, softmax_scale, causal, q_pos0, kv_pos0) -> (o, lse)`` is the building block
of tree attention as a first-class PyTorch operator: it has a fake (meta) implementation for shape inference /
tracing, and an autograd formula that differentiates through BOTH outputs -- ``lse`` carries gradient too, which is
what makes a user-written combine of several partials (``out = sum_r o_r * exp(lse_r - lse)``) differentiable end to
end.  Kernels: the same dispatch as ``ops.local.attention_partial`` (sm_100a decode / tcgen05 forward, tcgen05
backward; PyTorch oracle on CPU).  The reference has no operator registration and no backward at all."""
from __future__ import annotations

from typing import Tuple

import torch

from . import local as local_ops
from . import reference as ref


@torch.library.custom_op("tree_attention::partial", mutates_args=())
def partial(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, softmax_scale: float, causal: bool, q_pos0: int,
            kv_pos0: int) -> Tuple[torch.Tensor, torch.Tensor]:
    o, lse = local_ops.attention_partial(q, k, v, softmax_scale, causal, q_pos0, kv_pos0)
    return o.to(q.dtype).contiguous(), lse.float().contiguous()


@partial.register_fake
def _(q, k, v, softmax_scale, causal, q_pos0, kv_pos0):
    return q.new_empty(q.shape), q.new_empty(q.shape[:-1], dtype=torch.float32)


def _setup_context(ctx, inputs, output):
    q, k, v, softmax_scale, causal, q_pos0, kv_pos0 = inputs
    o, lse = output
    ctx.save_for_backward(q, k, v, o, lse)
    ctx.meta = (float(softmax_scale), bool(causal), int(q_pos0), int(kv_pos0))


def _backward(ctx, do, dlse):
    from .autograd import _local_bwd

    q, k, v, o, lse = ctx.saved_tensors
    scale, causal, q_pos0, kv_pos0 = ctx.meta
    do = torch.zeros_like(o) if do is None else do.contiguous()
    if dlse is not None:
        # d lse / d s_j = p_j: a cotangent on lse adds dS = P * dlse to the score gradient, i.e. dQ += dS K * scale and
        # dK += dS^T Q * scale (V is untouched).  P is recomputed blockwise from the saved lse.
        dq_e, dk_e = _lse_cotangent(q, k, lse, dlse.float(), scale, causal, q_pos0, kv_pos0)
    dq, dk, dv = _local_bwd(q, k, v, o, lse, do, scale, causal, q_pos0, kv_pos0)
    if dlse is not None:
        dq = dq.float() + dq_e
        dk = dk.float() + dk_e
    return dq.to(q.dtype), dk.to(k.dtype), dv.to(v.dtype), None, None, None, None


def _lse_cotangent(q, k, lse, dlse, scale, causal, q_pos0, kv_pos0, block: int = 4096):
    """Contribution of a cotangent on ``lse`` to (dq, dk): dS = P * dlse[..., None]."""
    b, hq, sq, d = q.shape
    hkv, s = k.shape[1], k.shape[2]
    g = hq // hkv
    qf = q.float().view(b, hkv, g, sq, d)
    dq = torch.zeros_like(qf)
    dk = torch.zeros(b, hkv, s, d, dtype=torch.float32, device=k.device)
    lse5 = lse.view(b, hkv, g, sq, 1)
    dl5 = dlse.view(b, hkv, g, sq, 1)
    rows = torch.arange(sq, device=q.device).view(1, 1, 1, sq, 1) + q_pos0
    for s0 in range(0, s, block):
        kb = k[:, :, s0:s0 + block].float().unsqueeze(2)                     # (B, Hkv, 1, blk, D)
        sc = torch.matmul(qf, kb.transpose(-1, -2)) * scale                  # (B, Hkv, G, Sq, blk)
        if causal:
            cols = torch.arange(s0, min(s0 + block, s), device=q.device).view(1, 1, 1, 1, -1) + kv_pos0
            sc = sc.masked_fill(cols > rows, float("-inf"))
        p = torch.exp(sc - lse5)
        p = torch.where(torch.isfinite(lse5), p, torch.zeros_like(p))
        ds = p * dl5
        dq += torch.matmul(ds, kb) * scale
        dk[:, :, s0:s0 + block] += torch.einsum("bhgqk,bhgqd->bhkd", ds, qf) * scale
    return dq.view(b, hq, sq, d), dk


torch.library.register_autograd("tree_attention::partial", _backward, setup_context=_setup_context)


def attention_partial_op(q, k, v, softmax_scale=None, causal=False, q_pos0=0, kv_pos0=0):
    """Differentiable ``(o, lse)`` through ``torch.ops.tree_attention.partial``."""
    scale = ref.default_scale(q.shape[-1]) if softmax_scale is None else float(softmax_scale)
    return torch.ops.tree_attention.partial(q, k, v, scale, bool(causal), int(q_pos0), int(kv_pos0))
