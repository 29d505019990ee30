"""Decode-time serving objects built on ``tree_attention``.

``TreeDecodeSession`` is the user-facing wrapper for the reference's one use case -- a decode step over a
sequence-sharded KV cache (``/root/reference/model.py:129-155`` runs exactly one such step) -- extended to
what a server needs: several KV caches ("layers") resident in HBM, a step captured once per layer in a
CUDA graph (the step is a single fused kernel; at 8 GPUs it is shorter than a Python launch), pinned-host
I/O for the query / result, and a KV-append for the owning rank.
"""
from __future__ import annotations

import time
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from ..parallel.tree import tree_attention


class TreeDecodeSession:
    def __init__(
        self,
        kv_layers: Sequence[Tuple[torch.Tensor, torch.Tensor]],
        softmax_scale: Optional[float] = None,
        causal: bool = False,
        backend: str = "fused",
        schedule: str = "oneshot",
        use_graph: bool = True,
        group=None,
        q_shape: Optional[Tuple[int, int, int, int]] = None,
        pdl: bool = False,
        dtype: Optional[torch.dtype] = None,
    ):
        """``pdl=True``: eager launches with programmatic dependent launch + K/V prefetch (the KV caches of this
        session are only written through ``append_kv``, never by the kernel that precedes a step) instead of CUDA
        graph replay -- the next step's prologue and first tile loads overlap the previous step's peer wait."""
        self.kv = list(kv_layers)
        k0 = self.kv[0][0]
        self.device = k0.device
        # quantised caches (MXFP8Tensor / MXFP8SeqTensor / FP8ChannelTensor) carry no query dtype: pass ``dtype``
        self.quantised = not isinstance(k0, torch.Tensor)
        self.dtype = dtype if dtype is not None else (torch.bfloat16 if self.quantised else k0.dtype)
        self.scale = softmax_scale
        self.causal = causal
        self.backend = backend
        self.schedule = schedule
        self.group = group
        self.pdl = 2 if pdl else 0
        self._kv_dirty = False  # set by append_kv: the next step must not prefetch K/V ahead of the append
        b, hkv, s, d = k0.shape
        self.q_shape = tuple(q_shape) if q_shape is not None else (b, hkv, 1, d)
        self.q_static = torch.zeros(self.q_shape, dtype=self.dtype, device=self.device)
        self.out_static: List[Optional[torch.Tensor]] = [None] * len(self.kv)
        self.graphs: List[torch.cuda.CUDAGraph] = []
        # end-to-end graphs: [H2D copy of the query from a pinned staging buffer | attention | D2H copy of the result]
        # as ONE graph launch per step (three runtime calls -> one on the latency path)
        self.e2e_graphs: List[torch.cuda.CUDAGraph] = []
        self.q_host: Optional[torch.Tensor] = None
        self.out_host: List[Optional[torch.Tensor]] = [None] * len(self.kv)
        g = self.q_shape[1] // hkv
        rows = g * self.q_shape[2]
        self.launches_per_step = 1
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        graphable = self.device.type == "cuda" and (world == 1 or backend in ("fused", "symm", "auto"))
        self._use_graph = bool(use_graph and graphable)
        self._prepared = False

    # -- internals ---------------------------------------------------------------------------------
    def _eager(self, q: torch.Tensor, layer: int, use_pdl: bool = False) -> torch.Tensor:
        k, v = self.kv[layer]
        pdl = 0
        if use_pdl:
            pdl = min(self.pdl, 1) if self._kv_dirty else self.pdl
        self._kv_dirty = False
        return tree_attention(q, k, v, group=self.group, causal=self.causal, softmax_scale=self.scale,
                              backend=self.backend, schedule=self.schedule, decode_pdl=pdl)

    def _prepare(self) -> None:
        if self._prepared:
            return
        for i in range(len(self.kv)):  # allocates workspaces and (collectively) the symmetric region
            self.out_static[i] = self._eager(self.q_static, i)
        if self.device.type == "cuda":
            torch.cuda.synchronize()
        if self._use_graph:
            side = torch.cuda.Stream()
            for i in range(len(self.kv)):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.stream(side):
                    with torch.cuda.graph(g):
                        self.out_static[i] = self._eager(self.q_static, i)
                self.graphs.append(g)
            torch.cuda.synchronize()
            self.q_host = torch.empty(self.q_shape, dtype=self.dtype).pin_memory()
            self.q_host.copy_(self.q_static.cpu())
            for i in range(len(self.kv)):
                self.out_host[i] = torch.empty(self.q_shape, dtype=self.dtype).pin_memory()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.stream(side):
                    with torch.cuda.graph(g):
                        self.q_static.copy_(self.q_host, non_blocking=True)
                        o = self._eager(self.q_static, i)
                        self.out_host[i].copy_(o, non_blocking=True)
                self.e2e_graphs.append(g)
            torch.cuda.synchronize()
        self._prepared = True

    # -- API ---------------------------------------------------------------------------------------
    def step_device(self, q: Optional[torch.Tensor], layer: int) -> torch.Tensor:
        """One decode-attention step for ``layer`` with the query already on the device.
        ``q=None`` reuses whatever is in ``q_static``."""
        layer %= len(self.kv)
        self._prepare()
        if q is not None and q.data_ptr() != self.q_static.data_ptr():
            self.q_static.copy_(q, non_blocking=True)
        if self.pdl:  # throughput path: eager launches chained by programmatic dependent launch
            return self._eager(self.q_static, layer, use_pdl=True)
        if self.graphs:
            self.graphs[layer].replay()
            return self.out_static[layer]
        self.out_static[layer] = self._eager(self.q_static, layer)
        return self.out_static[layer]

    def step(self, q_host: torch.Tensor, out_host: torch.Tensor, layer: int) -> torch.Tensor:
        """End-to-end (latency) step: pinned host query -> device, attention, result -> pinned host, synchronised.
        Uses the CUDA-graph replay when one was captured (lowest host overhead for a lone step)."""
        layer %= len(self.kv)
        self._prepare()
        if self.e2e_graphs:
            # the query goes through the session's pinned staging buffer (a host memcpy of a few KB); the graph's first
            # node copies it to the device, its last node copies the result into pinned host memory
            if q_host.data_ptr() != self.q_host.data_ptr():
                self.q_host.copy_(q_host)
            self.e2e_graphs[layer].replay()
            torch.cuda.current_stream().synchronize()
            if out_host.data_ptr() != self.out_host[layer].data_ptr():
                out_host.copy_(self.out_host[layer])
            return out_host
        self.q_static.copy_(q_host, non_blocking=True)
        if self.graphs:
            self.graphs[layer].replay()
            out = self.out_static[layer]
        else:
            out = self._eager(self.q_static, layer)
        out_host.copy_(out, non_blocking=True)
        if self.device.type == "cuda":
            torch.cuda.current_stream().synchronize()
        return out_host

    def append_kv(self, layer: int, k_new: torch.Tensor, v_new: torch.Tensor, position: int) -> None:
        """Write one new token's K/V at local row ``position`` of this rank's shard (the owner of the
        global position calls this; shards are preallocated)."""
        k, v = self.kv[layer]
        self._kv_dirty = True
        for cache, new in ((k, k_new), (v, v_new)):
            if hasattr(cache, "write_rows"):      # quantised cache: quantise on the way in
                cache.write_rows(position, new)
            else:
                cache[:, :, position : position + new.shape[2]].copy_(new, non_blocking=True)

    def run_e2e(self, q: torch.Tensor, steps: int, barrier) -> dict:
        self._prepare()
        qh = q.detach().cpu().pin_memory()
        oh = torch.empty(self.q_shape, dtype=self.dtype).pin_memory()
        for i in range(3):
            self.step(qh, oh, i)
        barrier()
        t0 = time.perf_counter()
        for i in range(steps):
            self.step(qh, oh, i)
        t1 = time.perf_counter()
        barrier()
        return {"ms": (t1 - t0) * 1e3, "h2d": qh.numel() * qh.element_size(), "d2h": oh.numel() * oh.element_size()}
