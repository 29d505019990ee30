"""Symmetric-memory runtime: the data plane of the fused kernels.

The reference's data plane is three blocking NCCL all-reduces (``/root/reference/model.py:108,114,115``).
Here every rank owns one device buffer per *kernel family* that is mapped into the address space of
every peer on the node (CUDA IPC, NVLink P2P through the NVSwitch), and the attention kernels
themselves store partials into the peers' buffers, release epoch flags with ``st.release.sys`` and
acquire the peers' flags with ``ld.acquire.sys`` (``csrc/decode_simt.cu``, ``csrc/combine.cu``,
``csrc/attn_fwd_sm100.cu``).  ``torch.distributed`` is used only to exchange the 64-byte IPC handles.

Buffer layout (per family, per rank)::

    [0      ) epoch   u32   device-resident call counter (bumped by the kernel => CUDA-graph safe); the fused decode
                            kernels use [0, 8) as a 64-bit arrival counter instead (4096 per launch)
    [64     ) status  16xu32  [0]=error code [1]=item [2]=source rank [3]=epoch, [8]=done counter
    [4096   ) flags   u32[flag_bytes/4]
    [4096+flag_bytes, ...) data  (float)

Families are separate allocations with separate epochs so that two kernels with different slot
layouts can never overwrite each other's in-flight data ("epoch protocol").

Allocation is collective: every rank of the group must request the same family in the same order.
Two providers: ``ipc`` (default; ``cudaMalloc`` + ``cudaIpc*`` in ``csrc/bindings.cpp``) and ``torch``
(``torch.distributed._symmetric_memory``; select with ``TREE_ATTN_SYMM=torch``).
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

from .. import _build
from ..utils.logging import logger

HEADER_BYTES = 4096
STATUS_OFFSET = 64


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def _next_pow2(x: int) -> int:
    return 1 << max(0, (int(x) - 1).bit_length())


@dataclass
class SymmRegion:
    family: str
    group: object
    rank: int
    world: int
    flag_bytes: int
    data_bytes: int
    total_bytes: int
    provider: str
    local_ptr: int
    peer_ptrs: List[int]
    comm: object  # _C.Comm
    keepalive: list = field(default_factory=list)
    layout: object = None  # slot layout of the last launch (see get_region)

    def status(self) -> Tuple[int, int, int, int]:
        C = _build.load()
        v = C.symm_read_u32(self.local_ptr + STATUS_OFFSET, 4)
        return tuple(int(x) for x in v)

    def epoch(self) -> int:
        """Launches of this family so far.  The fused decode kernels keep a 64-bit ARRIVAL counter at offset 0 (every
        launch advances it by 4096, csrc/decode_comm.cuh); the other families keep a 32-bit epoch there."""
        C = _build.load()
        lo, hi = (int(x) for x in C.symm_read_u32(self.local_ptr, 2))
        if self.family.startswith("decode"):
            return ((hi << 32) | lo) >> 12
        return lo

    def check(self) -> None:
        """Raise if a kernel of this family reported a bounded-spin timeout (failure detection)."""
        code, item, src, ep = self.status()
        if code != 0:
            raise RuntimeError(
                f"[tree_attention] family {self.family!r}: rank {src} never arrived at epoch {ep} "
                f"(item {item}) -- peer dead, not launched, or publish skipped"
            )

    def combine_stamps(self, reset: bool = True) -> dict:
        """In-kernel globaltimer stamps of the fused combine (max over CTAs and calls since the last reset):
        wait for peers, publish -> merged output written, CTA start -> local partial published (ns)."""
        C = _build.load()
        v = C.symm_read_u32(self.local_ptr + STATUS_OFFSET + 40, 3)
        if reset:
            torch.cuda.synchronize()
            C.symm_memset(self.local_ptr + STATUS_OFFSET + 40, 0, 12)
            torch.cuda.synchronize()
        return {"wait_peers_ns": int(v[0]), "combine_step_ns": int(v[1]), "local_partial_ns": int(v[2])}

    def clear_status(self) -> None:
        C = _build.load()
        C.symm_memset(self.local_ptr + STATUS_OFFSET, 0, 16)


_REGIONS: Dict[Tuple[str, int], SymmRegion] = {}


def _group_key(group) -> int:
    return 0 if group is None else id(group)


def _provider() -> str:
    return os.environ.get("TREE_ATTN_SYMM", "ipc")


def _alloc_ipc(total: int, group) -> Tuple[int, List[int], list]:
    C = _build.load()
    ptr, handle = C.symm_alloc(total)
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    handles: List[Optional[bytes]] = [None] * world
    dist.all_gather_object(handles, bytes(handle), group=group)
    ptrs = []
    for r, h in enumerate(handles):
        ptrs.append(ptr if r == rank else C.symm_open(h))
    dist.barrier(group)
    return ptr, ptrs, []


def _alloc_torch(total: int, group) -> Tuple[int, List[int], list]:
    import torch.distributed._symmetric_memory as symm_mem

    g = group if group is not None else dist.group.WORLD
    t = symm_mem.empty(total, dtype=torch.uint8, device=torch.device("cuda", torch.cuda.current_device()))
    t.zero_()
    hdl = symm_mem.rendezvous(t, g.group_name)
    ptrs = [int(p) for p in hdl.buffer_ptrs]
    torch.cuda.synchronize()
    dist.barrier(group)
    return int(t.data_ptr()), ptrs, [t, hdl]


def get_region(
    family: str,
    data_bytes: int,
    flag_bytes: int,
    group=None,
    timeout_s: Optional[float] = None,
    layout=None,
) -> SymmRegion:
    """Return (allocating or growing collectively if needed) the symmetric region of ``family``.

    ``layout`` identifies how the next kernel will carve the region into slots (e.g. the problem shape).  The
    parity double-buffering argument assumes consecutive launches of a family use the SAME carving; when the
    layout changes, a slow rank could still be reading the previous launch's slots under the old carving, so the
    change is fenced with a device sync + group barrier (only on shape changes, never in a steady-state loop)."""
    assert dist.is_initialized(), "symmetric memory needs an initialised process group"
    key = (family, _group_key(group))
    reg = _REGIONS.get(key)
    if reg is not None and reg.data_bytes >= data_bytes and reg.flag_bytes >= flag_bytes:
        if layout is not None and reg.layout is not None and layout != reg.layout:
            torch.cuda.synchronize()
            dist.barrier(group)
        if layout is not None:
            reg.layout = layout
        return reg
    if reg is not None:
        release(family, group)
    C = _build.load()
    flag_bytes = max(_round_up(_next_pow2(max(flag_bytes, 1)), 4096), 4096)
    data_bytes = max(_round_up(_next_pow2(max(data_bytes, 1)), 4096), 1 << 20)
    total = HEADER_BYTES + flag_bytes + data_bytes
    provider = _provider()
    torch.cuda.synchronize()
    if provider == "torch":
        ptr, ptrs, keep = _alloc_torch(total, group)
    else:
        ptr, ptrs, keep = _alloc_ipc(total, group)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    tmo = float(timeout_s if timeout_s is not None else os.environ.get("TREE_ATTN_COMM_TIMEOUT_S", "20"))
    comm = C.Comm(
        rank, world,
        [p + HEADER_BYTES + flag_bytes for p in ptrs],
        [p + HEADER_BYTES for p in ptrs],
        ptr, ptr + STATUS_OFFSET, data_bytes, flag_bytes, tmo,
    )
    reg = SymmRegion(family, group, rank, world, flag_bytes, data_bytes, total, provider, ptr, ptrs, comm, keep, layout)
    _REGIONS[key] = reg
    logger.debug(f"symm region {family!r}: {total >> 10} KiB per rank via {provider}")
    return reg


def release(family: str, group=None) -> None:
    key = (family, _group_key(group))
    reg = _REGIONS.pop(key, None)
    if reg is None:
        return
    torch.cuda.synchronize()
    if dist.is_initialized():
        try:
            dist.barrier(reg.group)
        except Exception:  # pragma: no cover - teardown best effort
            pass
    if reg.provider == "ipc":
        C = _build.load()
        for r, p in enumerate(reg.peer_ptrs):
            if r != reg.rank:
                try:
                    C.symm_close(p)
                except Exception:  # pragma: no cover
                    pass
        if dist.is_initialized():
            try:
                dist.barrier(reg.group)
            except Exception:  # pragma: no cover
                pass
        C.symm_free(reg.local_ptr)
    reg.keepalive.clear()


def release_all() -> None:
    for fam, gk in list(_REGIONS.keys()):
        reg = _REGIONS.get((fam, gk))
        if reg is not None:
            release(fam, reg.group)


def regions() -> Dict[Tuple[str, int], SymmRegion]:
    return _REGIONS
