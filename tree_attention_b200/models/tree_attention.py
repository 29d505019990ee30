"""``nn.Module`` front-ends.

The reference has no modules or parameters ("no nn.Module, no requires_grad"; the op itself is
``/root/reference/model.py:85-124``).  A user
switching frameworks still needs the op as a layer, so:

* ``TreeAttention``      -- the bare op as a module (replicated q, sequence-sharded k/v).
* ``TreeSelfAttention``  -- fused QKV projection + tree attention + output projection over a sequence-sharded
                            hidden state: each rank projects its own tokens to K/V (its shard), queries come
                            from the (replicated) query tokens.  Used by the smoke test and the examples.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from ..parallel.tree import tree_attention


class TreeAttention(nn.Module):
    def __init__(self, causal: bool = False, softmax_scale: Optional[float] = None, backend: str = "auto",
                 schedule: str = "oneshot", group=None, layout: str = "bhsd", kv_layout: str = "contiguous"):
        super().__init__()
        self.causal, self.softmax_scale, self.backend, self.schedule = causal, softmax_scale, backend, schedule
        self.group, self.layout, self.kv_layout = group, layout, kv_layout   # kv_layout="zigzag": balanced causal shards

    def forward(self, q, k, v, kv_offset: Optional[int] = None, q_offset: Optional[int] = None):
        if torch.is_grad_enabled() and (q.requires_grad or k.requires_grad or v.requires_grad):
            from ..ops.autograd import tree_attention_func

            return tree_attention_func(q, k, v, causal=self.causal, softmax_scale=self.softmax_scale,
                                       group=self.group, kv_offset=kv_offset, q_offset=q_offset, backend=self.backend,
                                       schedule=self.schedule, layout=self.layout, kv_layout=self.kv_layout)
        return tree_attention(q, k, v, group=self.group, causal=self.causal, softmax_scale=self.softmax_scale,
                              kv_offset=kv_offset, q_offset=q_offset, backend=self.backend, schedule=self.schedule,
                              layout=self.layout, kv_layout=self.kv_layout)


class TreeSelfAttention(nn.Module):
    """x_q: (B, Sq, E) replicated query tokens; x_kv: (B, S_local, E) this rank's slice of the context."""

    def __init__(self, embed_dim: int, num_heads: int, num_kv_heads: Optional[int] = None, causal: bool = False,
                 bias: bool = False, group=None, dtype=None, device=None, kv_layout: str = "contiguous"):
        super().__init__()
        self.h, self.hkv = num_heads, num_kv_heads or num_heads
        assert embed_dim % num_heads == 0 and num_heads % self.hkv == 0
        self.d = embed_dim // num_heads
        kw = dict(dtype=dtype, device=device)
        self.q_proj = nn.Linear(embed_dim, self.h * self.d, bias=bias, **kw)
        self.kv_proj = nn.Linear(embed_dim, 2 * self.hkv * self.d, bias=bias, **kw)
        self.o_proj = nn.Linear(self.h * self.d, embed_dim, bias=bias, **kw)
        self.attn = TreeAttention(causal=causal, group=group, kv_layout=kv_layout)   # "zigzag": x_kv = zigzag_shard(tokens)

    def forward(self, x_q: torch.Tensor, x_kv: torch.Tensor, kv_offset: Optional[int] = None,
                q_offset: Optional[int] = None) -> torch.Tensor:
        b, sq, _ = x_q.shape
        s = x_kv.shape[1]
        q = self.q_proj(x_q).view(b, sq, self.h, self.d).transpose(1, 2)
        kv = self.kv_proj(x_kv).view(b, s, 2, self.hkv, self.d)
        k, v = kv[:, :, 0].transpose(1, 2), kv[:, :, 1].transpose(1, 2)
        o = self.attn(q, k, v, kv_offset=kv_offset, q_offset=q_offset)
        return self.o_proj(o.transpose(1, 2).reshape(b, sq, self.h * self.d))
