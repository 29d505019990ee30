"""``python3 model.py`` -- the reference's entry point, on the B200-native framework.

Same UX and the same public names as ``/root/reference/model.py`` (``setup``, ``cleanup``, ``make_data``,
``flash_res_lse``, ``tree_decode``, ``main`` and the zero-argument CLI, model.py:159-169): one process is
spawned per visible GPU (or a single CPU rank), each rank builds its KV shard of the reference's default
problem (B=1, 16 heads, T=64000 per rank, C=128, fp16, model.py:140-145), runs one tree-decode step and
logs the time.  What differs:

* the step is ONE fused sm_100a kernel per rank (no NCCL on the hot path) and it actually runs for
  world_size > 1 (the reference raises at model.py:111);
* timing is CUDA events after warm-up, max over ranks (D10); the output is validated against a
  float64 oracle assembled from all shards (D12); nothing logs inside the timed region (D14);
* every literal is a flag / env var (``--help``), defaults unchanged (5.6); torchrun is supported
  (``RANK``/``WORLD_SIZE``/``MASTER_*``); ``--json`` prints one machine-readable line.
"""
from __future__ import annotations

import json
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tree_attention_b200 import (  # noqa: F401  (re-exported: the reference's public names)
    cleanup,
    flash_res_lse,
    make_data,
    setup,
    tree_attention,
    tree_decode,
)
from tree_attention_b200.ops import reference as ref
from tree_attention_b200.utils.config import TreeAttentionConfig, from_args
from tree_attention_b200.utils.logging import add_file_sink, logger
from tree_attention_b200.utils.timing import time_cuda, time_host

_CFG: TreeAttentionConfig = TreeAttentionConfig()


def _oracle_check(cfg: TreeAttentionConfig, q, k, v, out, rank: int, world_size: int) -> float:
    """Gather every shard's partial oracle on all ranks and compare (float64 merge of fp32 partials)."""
    scale = 1.0 if cfg.softmax_scale is None else cfg.softmax_scale
    s_local = k.shape[2]
    q_pos0 = world_size * s_local - q.shape[2]
    o_p, l_p = ref.attention_partial_ref(q, k, v, scale, cfg.causal, q_pos0, rank * s_local, torch.float32,
                                         block=16384)
    if world_size > 1:
        packed = torch.cat([o_p.float(), l_p[..., None]], dim=-1).contiguous()
        bufs = [torch.empty_like(packed) for _ in range(world_size)]
        dist.all_gather(bufs, packed)
        o_ref, _ = ref.merge_many([b[..., :-1].double() for b in bufs], [b[..., -1].double() for b in bufs])
    else:
        o_ref = o_p.double()
    return float((out.double() - o_ref).abs().max())


def main(rank: int, world_size: int) -> None:
    """Per-rank driver (model.py:129): init, data, one timed tree-decode step, validate, clean up."""
    cfg = _CFG
    device = torch.device(f"cuda:{rank}" if torch.cuda.is_available() else "cpu")
    # Under torchrun the rendezvous belongs to the launcher: the workers must connect to ITS store (MASTER_ADDR/PORT of the
    # environment; TORCHELASTIC_USE_AGENT_STORE means rank 0 does not open one of its own), so the config's default port
    # (the reference's 12355, model.py:21) only applies to the self-spawned path.  Overriding it hung all 8 ranks (round 2).
    under_launcher = "TORCHELASTIC_RUN_ID" in os.environ or ("RANK" in os.environ and "MASTER_PORT" in os.environ)
    setup(rank, world_size,
          master_addr=None if under_launcher else cfg.master_addr,
          master_port=None if under_launcher else cfg.master_port,
          local_rank=int(os.environ["LOCAL_RANK"]) if "LOCAL_RANK" in os.environ else None)
    dtype = cfg.torch_dtype if device.type == "cuda" else torch.float32
    shape = (cfg.batch, cfg.num_heads, cfg.seq_len, cfg.head_dim)
    qfl, kfl, vfl = make_data(shape, rank, device, dtype=dtype, layout=cfg.layout, sq=cfg.q_len,
                              num_kv_heads=cfg.num_kv_heads, seed=cfg.seed)
    logger.info(f"Rank {rank}: Starting computation with seq_len: {cfg.seq_len}, hid_dim: {cfg.num_heads * cfg.head_dim}")
    scale = 1.0 if cfg.softmax_scale is None else cfg.softmax_scale  # the reference's default (model.py:60,100)
    k_ref, v_ref = kfl, vfl
    if cfg.kv_format != "native":   # quantised KV cache; the oracle check runs on the de-quantised cache
        from tree_attention_b200.ops.quant import FP8ChannelTensor, MXFP8SeqTensor, MXFP8Tensor

        if cfg.layout != "bhsd":
            raise SystemExit("--kv-format needs --layout bhsd")
        if cfg.kv_format == "fp8":
            kfl, vfl = FP8ChannelTensor.from_float(kfl), FP8ChannelTensor.from_float(vfl)
        elif cfg.kv_format == "mxfp8":
            kfl, vfl = MXFP8Tensor.from_float(kfl), MXFP8SeqTensor.from_float(vfl)
        elif cfg.kv_format == "mxfp8-simt":
            kfl, vfl = MXFP8Tensor.from_float(kfl), MXFP8Tensor.from_float(vfl)
        else:
            raise SystemExit(f"unknown --kv-format {cfg.kv_format}")
        k_ref, v_ref = kfl.dequantize(dtype), vfl.dequantize(dtype)

    def step():
        return tree_attention(qfl, kfl, vfl, causal=cfg.causal, softmax_scale=scale, backend=cfg.backend,
                              schedule=cfg.schedule)

    warmup = cfg.warmup if cfg.warmup > 0 else (2 if device.type == "cuda" else 0)
    if device.type == "cuda":
        t = time_cuda(step, steps=max(cfg.steps, 1), warmup=warmup)
        seconds = t["ms_per_step"] * 1e-3
    else:
        t = time_host(step, steps=max(cfg.steps, 1), warmup=warmup)
        seconds = t["ms_per_step"] * 1e-3
    output = step()
    logger.info(f"Rank {rank}: Computation completed in {seconds}s")
    err = None
    if cfg.check:
        err = _oracle_check(cfg, qfl, k_ref, v_ref, output, rank, world_size)
        tol = 2e-2 if dtype in (torch.float16, torch.bfloat16) else 1e-4
        if cfg.kv_format != "native":
            tol = 8e-2   # q and P are rounded to e4m3 inside the fp8 kernels
        status = "OK" if err < tol else "MISMATCH"
        logger.info(f"Rank {rank}: max |out - oracle| = {err:.3e} [{status}]")
        if err >= tol:
            cleanup()
            raise SystemExit(f"rank {rank}: output mismatch vs oracle ({err:.3e})")
    if cfg.json and rank == 0:
        s_global = cfg.seq_len * world_size
        print(json.dumps({
            "world_size": world_size, "device": device.type, "dtype": str(dtype).replace("torch.", ""),
            "shape": {"B": cfg.batch, "Hq": cfg.num_heads, "Hkv": cfg.kv_heads, "Sq": cfg.q_len,
                      "S_per_rank": cfg.seq_len, "S_global": s_global, "D": cfg.head_dim},
            "steps": max(cfg.steps, 1), "warmup": warmup, "timer": "cuda_events" if device.type == "cuda" else "host_clock",
            "latency_us": seconds * 1e6, "decode_tokens_per_s": cfg.batch * cfg.q_len / seconds,
            "kv_tokens_per_s": cfg.batch * s_global / seconds, "max_abs_err": err,
            "backend": cfg.backend, "schedule": cfg.schedule, "kv_format": cfg.kv_format,
        }))
    cleanup()


def _spawn_entry(rank: int, world_size: int, cfg: TreeAttentionConfig) -> None:
    global _CFG
    _CFG = cfg
    main(rank, world_size)


def cli() -> None:
    """`python3 model.py` / the `tree-attention` console script: one rank per visible GPU, or one CPU rank
    (reference launcher: /root/reference/model.py:159-169)."""
    global _CFG
    _CFG = from_args()
    add_file_sink(_CFG.log_file, _CFG.log_rotation)  # model.py:160
    if "RANK" in os.environ and "WORLD_SIZE" in os.environ:  # launched by torchrun
        main(int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]))
    elif torch.cuda.is_available():
        world_size = torch.cuda.device_count()
        logger.info(f"Running on {world_size} GPUs.")
        if world_size > 1:
            mp.spawn(_spawn_entry, args=(world_size, _CFG), nprocs=world_size, join=True)
        else:
            main(0, 1)
    else:
        logger.info("Running on CPU.")
        main(rank=0, world_size=1)


if __name__ == "__main__":
    cli()
    sys.exit(0)
