"""Process-group bootstrap: one process per GPU, ``torch.distributed`` for the plumbing.

Reference parity: ``setup`` / ``cleanup`` (``/root/reference/model.py:11-33``).  The reference
hard-codes backend ``"nccl"``, ``localhost:12355`` and silently does nothing on CPU
(model.py:19-22), so a CPU run can never be multi-rank.  Here:

* backend is NCCL when CUDA is present, gloo otherwise (override with ``backend=`` or
  ``TREE_ATTN_BACKEND``);
* rendezvous address/port come from arguments, then ``MASTER_ADDR``/``MASTER_PORT`` (torchrun),
  then ``127.0.0.1:12355`` (the reference's port, with a resolvable address);
* ``torch.cuda.set_device(local_rank)`` is called before the group is created, which the
  reference never does;
* ``cleanup`` is symmetric with ``setup`` on every backend.

NCCL is only the bootstrap and the *baseline* data path.  The product data path is the
symmetric-memory runtime in ``parallel/symm.py``.
"""
from __future__ import annotations

import datetime
import os
from dataclasses import dataclass
from typing import Optional

import torch
import torch.distributed as dist

from ..utils.logging import logger

DEFAULT_ADDR = "127.0.0.1"
DEFAULT_PORT = 12355  # model.py:21


@dataclass
class RuntimeState:
    rank: int = 0
    world_size: int = 1
    local_rank: int = 0
    backend: str = "none"
    device: torch.device = torch.device("cpu")
    owns_group: bool = False
    initialised: bool = False


_STATE = RuntimeState()


def get_runtime() -> RuntimeState:
    return _STATE


def pick_backend(backend: Optional[str] = None) -> str:
    backend = backend or os.environ.get("TREE_ATTN_BACKEND")
    if backend:
        return backend
    return "nccl" if torch.cuda.is_available() else "gloo"


def env_rank_world() -> tuple[int, int, int]:
    """(rank, world_size, local_rank) from torchrun-style environment, defaults (0, 1, 0)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    return rank, world, local


def setup(
    rank: int,
    world_size: int,
    backend: Optional[str] = None,
    master_addr: Optional[str] = None,
    master_port: Optional[int] = None,
    local_rank: Optional[int] = None,
    timeout_s: float = 600.0,
) -> RuntimeState:
    """Initialise the distributed environment (model.py:11).  Works on CPU (gloo) and GPU (NCCL).

    ``world_size == 1`` needs no process group; the call still records rank/device so that the
    rest of the API behaves uniformly.
    """
    global _STATE
    backend = pick_backend(backend)
    local_rank = rank if local_rank is None else local_rank
    if torch.cuda.is_available():
        ndev = torch.cuda.device_count()
        device = torch.device("cuda", local_rank % max(ndev, 1))
        torch.cuda.set_device(device)
    else:
        device = torch.device("cpu")
    owns = False
    if world_size > 1 and not dist.is_initialized():
        addr = master_addr or os.environ.get("MASTER_ADDR") or DEFAULT_ADDR
        if addr == "localhost":
            addr = DEFAULT_ADDR  # the container hostname/localhost may not resolve
        port = int(master_port or os.environ.get("MASTER_PORT") or DEFAULT_PORT)
        os.environ["MASTER_ADDR"] = addr
        os.environ["MASTER_PORT"] = str(port)
        kwargs = {}
        if backend == "nccl":
            kwargs["device_id"] = device
        dist.init_process_group(
            backend,
            rank=rank,
            world_size=world_size,
            timeout=datetime.timedelta(seconds=timeout_s),
            **kwargs,
        )
        owns = True
    elif world_size > 1:
        backend = dist.get_backend()
    _STATE = RuntimeState(
        rank=rank,
        world_size=world_size,
        local_rank=local_rank,
        backend=backend if world_size > 1 else "none",
        device=device,
        owns_group=owns,
        initialised=True,
    )
    logger.info(
        f"Distributed environment initialized on rank {rank} with {world_size} "
        f"{'GPUs' if device.type == 'cuda' else 'CPU ranks'} (backend={_STATE.backend})."
    )
    return _STATE


def cleanup() -> None:
    """Tear down what ``setup`` created (model.py:27).  Safe to call when nothing was set up."""
    global _STATE
    from . import symm

    symm.release_all()
    if _STATE.owns_group and dist.is_initialized():
        dist.destroy_process_group()
        logger.info("Distributed environment cleaned up.")
    _STATE = RuntimeState()
