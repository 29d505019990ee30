"""Decode-time serving objects built on ``tree_attention``.

``TreeDecodeSession`` is the user-facing wrapper for the reference's one use case -- a decode step over a
sequence-sharded KV cache (``/root/reference/model.py:129-155`` runs exactly one such step) -- extended to
what a server needs:

* several KV caches ("layers") resident in HBM, each with a device-resident FILL LEVEL (``kv_len``): the kernels
  read it at run time, so rows that have not been written yet never enter the softmax and a CUDA graph captured
  once follows the growing cache (``append_kv`` bumps the level on the device);
* a step prepared ONCE per layer in C++ (``_C.DecodeStep``: tensor maps encoded, parameter block filled) and
  re-launched with a single runtime call, optionally chained by programmatic dependent launch, or replayed from a
  CUDA graph; the session OWNS the split-merge workspace and its symmetric region, so nothing a graph or a prepared
  launch points at can be re-allocated behind its back;
* pinned-host I/O for the query / result (``step``).
"""
from __future__ import annotations

import itertools
import time
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from ..ops import local as local_ops
from ..ops import reference as ref
from ..parallel import symm
from ..parallel.tree import tree_attention

_SESSION_IDS = itertools.count()


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
        kv_lens: Optional[Sequence[int]] = None,
        host_io: str = "copy",
    ):
        """``kv_layers``: per layer this rank's preallocated ``(k, v)`` shard, ``(B, Hkv, capacity, D)``.
        ``kv_lens``: per layer the number of rows already filled (default: the shards are full).  Construction is
        collective (every rank of ``group`` builds its session in the same order).
        ``pdl=True``: ``step_device`` launches eagerly with programmatic dependent launch + K/V prefetch (the caches of
        a session are only written through ``append_kv``) -- the next step's prologue and first tile loads overlap
        the previous step's drain; the e2e path (``step``) replays a CUDA graph.
        ``host_io``: how the latency path (``step``) moves its few KB across PCIe.  ``"copy"`` (default): a CUDA graph
        [H2D memcpy | attention | D2H memcpy].  ``"zero_copy"`` (native fast path only): the decode kernel itself loads q
        from the session's pinned, device-mapped staging buffer and stores the result straight into pinned host memory --
        ONE kernel launch per step, no copy-engine hops.  Measured equal within noise on one B200 at 128K (0.342 vs
        0.341 ms per step, profiles/r2_decode/e2e_host_io.md): the PCIe read of q inside the kernel costs what the two
        memcpy nodes cost, the rest of the end-to-end overhead is the lone launch + stream sync."""
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
        self._kv_dirty = False  # set by append_kv: the next step must not prefetch K/V (or read kv_len) ahead of the append
        b, hkv, s, d = k0.shape
        self.capacity = s
        self.q_shape = tuple(q_shape) if q_shape is not None else (b, hkv, 1, d)
        self.q_static = torch.zeros(self.q_shape, dtype=self.dtype, device=self.device)
        self.out_static: List[Optional[torch.Tensor]] = [None] * len(self.kv)
        self.graphs: List[torch.cuda.CUDAGraph] = []
        # end-to-end graphs: [H2D copy of the query from a pinned staging buffer | attention | D2H copy of the result]
        # as ONE graph launch per step (three runtime calls -> one on the latency path)
        self.e2e_graphs: List[torch.cuda.CUDAGraph] = []
        self.q_host: Optional[torch.Tensor] = None
        self.out_host: List[Optional[torch.Tensor]] = [None] * len(self.kv)
        self.launches_per_step = 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        graphable = self.device.type == "cuda" and (self.world == 1 or backend in ("fused", "symm", "auto"))
        self._use_graph = bool(use_graph and graphable)
        self._prepared = False
        # fill levels: host mirror + one int32 per layer on the device (the kernels' kv_len scalars are views of it)
        self.kv_len_host: Optional[List[int]] = None
        self.kv_len_dev: Optional[torch.Tensor] = None
        if kv_lens is not None:
            self.kv_len_host = [int(x) for x in kv_lens]
            assert len(self.kv_len_host) == len(self.kv) and all(0 <= x <= s for x in self.kv_len_host)
            self.kv_len_dev = torch.tensor(self.kv_len_host, dtype=torch.int32, device=self.device)
        # native fast path: plain bf16 / fp16 caches on CUDA through the fused (or single-GPU) decode kernels
        self._steps: List[object] = []
        self._steps_zc: List[object] = []   # zero-copy twins of _steps: q / out in pinned host memory
        assert host_io in ("zero_copy", "copy")
        self.host_io = host_io
        self._ws = None
        self.region = None
        self._family = None
        self._fast = (self.device.type == "cuda" and not self.quantised and backend in ("fused", "auto") and
                      schedule == "oneshot" and local_ops.decode_eligible(self.q_static, k0))

    @classmethod
    def allocate(cls, n_layers: int, b: int, hkv: int, capacity: int, d: int, device, dtype=torch.bfloat16, **kw):
        """Session over freshly allocated, ZERO-initialised caches with fill level 0 (rows past the fill level inside a
        128-row tile are multiplied by exact zeros on the tensor cores, so they must be finite -- zeros are)."""
        kv = [(torch.zeros(b, hkv, capacity, d, dtype=dtype, device=device),
               torch.zeros(b, hkv, capacity, d, dtype=dtype, device=device)) for _ in range(n_layers)]
        return cls(kv, kv_lens=[0] * n_layers, **kw)

    # -- internals ---------------------------------------------------------------------------------
    def _kv_len_of(self, layer: int):
        return None if self.kv_len_dev is None else self.kv_len_dev[layer : layer + 1]

    def _positions(self) -> Tuple[int, int]:
        """(q_pos0, kv_pos0) as tree_attention derives them: contiguous capacity-sized shards in rank order, the
        query at the end of the global sequence."""
        return self.world * self.capacity - self.q_shape[2], self.rank * self.capacity

    def _eager(self, q: torch.Tensor, layer: int, use_pdl: bool = False) -> torch.Tensor:
        k, v = self.kv[layer]
        pdl = 0
        if use_pdl:
            pdl = min(self.pdl, 1) if self._kv_dirty else self.pdl
        self._kv_dirty = False
        return tree_attention(q, k, v, group=self.group, causal=self.causal, softmax_scale=self.scale,
                              backend=self.backend, schedule=self.schedule, decode_pdl=pdl, kv_len=self._kv_len_of(layer))

    def _prepare_fast(self) -> None:
        from .. import _build

        C = _build.load()
        k0 = self.kv[0][0]
        b, hq, sq, d = self.q_shape
        impl = local_ops.decode_impl_for(self.q_shape, k0.shape)
        nfloats, ntick = local_ops.decode_workspace_sizes(self.q_shape, k0.shape, impl)
        self._ws = local_ops.new_workspace(self.device, nfloats, ntick)   # owned: never re-allocated under a graph
        comm = None
        if self.world > 1:
            data, flags = local_ops.decode_comm_bytes(b, hq, k0.shape[1], sq, self.capacity, d, self.world)
            self._family = f"decode.session{next(_SESSION_IDS)}"          # a region of its own, sized once
            self.region = symm.get_region(self._family, data, flags, self.group, layout=(b, hq, k0.shape[1], sq, d))
            comm = self.region.comm
        scale = ref.default_scale(d) if self.scale is None else float(self.scale)
        q_pos0, kv_pos0 = self._positions()
        for i, (k, v) in enumerate(self.kv):
            assert tuple(k.shape) == tuple(k0.shape), "all layers of a session share one shard shape"
            self.out_static[i] = torch.empty(self.q_shape, dtype=self.dtype, device=self.device)
            self._steps.append(C.decode_step(impl, self.q_static, k, v, self.out_static[i], None, self._ws["part"],
                                             self._ws["tickets"], comm, scale, bool(self.causal), int(q_pos0),
                                             int(kv_pos0), self._kv_len_of(i)))
            if self.host_io == "zero_copy":
                if self.q_host is None:
                    self.q_host = torch.zeros(self.q_shape, dtype=self.dtype).pin_memory()
                self.out_host[i] = torch.zeros(self.q_shape, dtype=self.dtype).pin_memory()
                # same workspace, same symmetric region, same launch-tag counters as the device-resident step: the two
                # are interchangeable launch by launch (every rank takes the same path for a given step)
                self._steps_zc.append(C.decode_step(impl, self.q_host, k, v, self.out_host[i], None, self._ws["part"],
                                                    self._ws["tickets"], comm, scale, bool(self.causal), int(q_pos0),
                                                    int(kv_pos0), self._kv_len_of(i)))
        self.launches_per_step = self._steps[0].kernels_per_step

    def _launch(self, layer: int, use_pdl: bool = False) -> torch.Tensor:
        if self._steps:
            pdl = 0
            if use_pdl:
                pdl = min(self.pdl, 1) if self._kv_dirty else self.pdl
            self._kv_dirty = False
            self._steps[layer].launch(pdl)
            return self.out_static[layer]
        return self._eager(self.q_static, layer, use_pdl)   # a fresh output tensor per call (graph capture pins it)

    def _prepare(self) -> None:
        if self._prepared:
            return
        if self._fast:
            self._prepare_fast()
        for i in range(len(self.kv)):  # first launches: allocates workspaces and (collectively) symmetric regions
            self.out_static[i] = self._launch(i)
        if self.device.type == "cuda":
            torch.cuda.synchronize()
        if self._use_graph:
            side = torch.cuda.Stream()
            for i in range(len(self.kv)):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.stream(side):
                    with torch.cuda.graph(g):
                        self.out_static[i] = self._launch(i)
                self.graphs.append(g)
            torch.cuda.synchronize()
            if self._steps_zc:          # zero-copy latency path: no e2e graphs needed
                for i in range(len(self.kv)):
                    self._steps_zc[i].launch(0)
                torch.cuda.synchronize()
                self._prepared = True
                return
            self.q_host = torch.empty(self.q_shape, dtype=self.dtype).pin_memory()
            self.q_host.copy_(self.q_static.cpu())
            for i in range(len(self.kv)):
                self.out_host[i] = torch.empty(self.q_shape, dtype=self.dtype).pin_memory()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.stream(side):
                    with torch.cuda.graph(g):
                        self.q_static.copy_(self.q_host, non_blocking=True)
                        o = self._launch(i)
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
        if self.pdl:  # throughput path: launches chained by programmatic dependent launch
            self.out_static[layer] = self._launch(layer, use_pdl=True)
            return self.out_static[layer]
        if self.graphs:
            self.graphs[layer].replay()
            return self.out_static[layer]
        self.out_static[layer] = self._launch(layer)
        return self.out_static[layer]

    def step(self, q_host: torch.Tensor, out_host: torch.Tensor, layer: int) -> torch.Tensor:
        """End-to-end (latency) step: pinned host query -> device, attention, result -> pinned host, synchronised.
        Uses the CUDA-graph replay when one was captured (lowest host overhead for a lone step)."""
        layer %= len(self.kv)
        self._prepare()
        if self._steps_zc:
            # zero-copy: the kernel reads q from the pinned staging buffer over PCIe and posts the result into pinned host
            # memory; the step is one prepared launch + a stream sync
            if q_host.data_ptr() != self.q_host.data_ptr():
                self.q_host.copy_(q_host)
            self._kv_dirty = False
            self._steps_zc[layer].launch(0)
            torch.cuda.current_stream().synchronize()
            if out_host.data_ptr() != self.out_host[layer].data_ptr():
                out_host.copy_(self.out_host[layer])
            return out_host
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
            out = self._launch(layer)
        out_host.copy_(out, non_blocking=True)
        if self.device.type == "cuda":
            torch.cuda.current_stream().synchronize()
        return out_host

    def append_kv(self, layer: int, k_new: torch.Tensor, v_new: torch.Tensor, position: Optional[int] = None) -> None:
        """Write new tokens' K/V at local row ``position`` of this rank's shard (default: at the current fill level) and
        raise the fill level to cover them.  The rank that owns the global position calls this; other ranks do not."""
        k, v = self.kv[layer]
        n = k_new.shape[2]
        if position is None:
            if self.kv_len_host is None:
                raise ValueError("append_kv without a position needs a session created with kv_lens")
            position = self.kv_len_host[layer]
        if position + n > self.capacity:
            raise ValueError(f"append_kv: rows {position}..{position + n} exceed the shard capacity {self.capacity}")
        self._kv_dirty = True
        for cache, new in ((k, k_new), (v, v_new)):
            if hasattr(cache, "write_rows"):      # quantised cache: quantise on the way in
                cache.write_rows(position, new)
            else:
                cache[:, :, position : position + n].copy_(new, non_blocking=True)
        if self.kv_len_host is not None and position + n > self.kv_len_host[layer]:
            self.kv_len_host[layer] = position + n
            # stream-ordered device update (a fill kernel: graph replays and prepared launches queued behind it see it)
            self.kv_len_dev[layer : layer + 1].fill_(position + n)

    def close(self) -> None:
        """Release the session's symmetric region (collective).  ``cleanup()`` does this for every live region."""
        self.graphs.clear()
        self.e2e_graphs.clear()
        self._steps.clear()
        self._steps_zc.clear()
        if self._family is not None:
            symm.release(self._family, self.group)
            self._family = None
            self.region = None

    def run_e2e(self, q: torch.Tensor, steps: int, barrier) -> dict:
        """Time ``steps`` end-to-end steps (host clock around the whole loop; per-step times are kept for diagnosis)."""
        self._prepare()
        if self.e2e_graphs or self._steps_zc:
            # the caller fills the session's pinned staging buffer and reads the per-layer pinned result buffers directly:
            # the step is then ONE graph launch [H2D copy | attention | D2H copy] + a stream sync, no extra host memcpy
            self.q_host.copy_(q.detach().cpu())
            qh, oh = self.q_host, None
        else:
            qh = q.detach().cpu().pin_memory()
            oh = torch.empty(self.q_shape, dtype=self.dtype).pin_memory()
        for i in range(3):
            self.step(qh, oh if oh is not None else self.out_host[i % len(self.kv)], i)
        barrier()
        per = []
        t0 = time.perf_counter()
        tp = t0
        for i in range(steps):
            self.step(qh, oh if oh is not None else self.out_host[i % len(self.kv)], i)
            tn = time.perf_counter()
            per.append((tn - tp) * 1e3)
            tp = tn
        t1 = time.perf_counter()
        barrier()
        per.sort()
        return {"ms": (t1 - t0) * 1e3, "h2d": qh.numel() * qh.element_size(), "d2h": qh.numel() * qh.element_size(),
                "median_ms": per[len(per) // 2], "max_ms": per[-1], "min_ms": per[0],
                "transfer": ("zero_copy: the decode kernel loads q from pinned mapped host memory and stores the result into "
                             "pinned host memory (1 launch + stream sync per step)") if self._steps_zc else
                            ("cuda_graph[H2D memcpy | attention | D2H memcpy] + stream sync" if self.e2e_graphs else
                             "eager H2D copy, launch, D2H copy, stream sync")}
