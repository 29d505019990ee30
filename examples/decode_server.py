"""Decode serving loop over a sequence-sharded (optionally fp8) KV cache.

    python examples/decode_server.py                          # one process per visible GPU, or one CPU rank
    python examples/decode_server.py --kv-format mxfp8 --steps 64
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 examples/decode_server.py

Every rank owns ``--tokens-per-rank`` keys of each layer's KV cache; the query of the new token is replicated.  A step
is ``TreeDecodeSession.step``: pinned-host query -> [H2D | fused attention + cross-GPU tree combine | D2H] replayed as
one CUDA graph -> pinned-host result (on CPU: the PyTorch oracle + gloo).  After each step the owner rank of the new
position appends the token's K/V (quantised on the way in for fp8 caches)."""
from __future__ import annotations

import argparse
import os
import sys
import time

import torch
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tree_attention_b200 as ta  # noqa: E402
from tree_attention_b200.models.decoder import TreeDecodeSession  # noqa: E402
from tree_attention_b200.ops.quant import FP8ChannelTensor, MXFP8SeqTensor, MXFP8Tensor  # noqa: E402


def worker(rank: int, world: int, a) -> None:
    dev = torch.device(f"cuda:{rank}" if torch.cuda.is_available() else "cpu")
    ta.setup(rank, world, master_port=a.port)
    dtype = torch.bfloat16 if dev.type == "cuda" else torch.float32
    cap = a.tokens_per_rank + a.steps          # room for the appended tokens on every rank
    layers = []
    for layer in range(a.layers):
        _, k, v = ta.make_data((1, a.heads, cap, a.head_dim), rank, dev, dtype=dtype, num_kv_heads=a.kv_heads,
                               seed=layer, log=False)
        k[:, :, a.tokens_per_rank:] = 0         # unwritten tail of the preallocated shard: excluded by the fill level
        v[:, :, a.tokens_per_rank:] = 0         # (kept finite: the tensor-core kernels multiply it by exact zeros)
        if a.kv_format == "fp8":
            k, v = FP8ChannelTensor.from_float(k, headroom=2.0), FP8ChannelTensor.from_float(v, headroom=2.0)
        elif a.kv_format == "mxfp8":
            k, v = MXFP8Tensor.from_float(k), MXFP8SeqTensor.from_float(v)
        layers.append((k, v))
    scale = a.head_dim ** -0.5
    sess = TreeDecodeSession(layers, softmax_scale=scale, q_shape=(1, a.heads, 1, a.head_dim), dtype=dtype,
                             backend="auto", kv_lens=[a.tokens_per_rank] * a.layers)
    g = torch.Generator().manual_seed(1234)    # the same queries / new tokens on every rank
    oh = torch.empty(1, a.heads, 1, a.head_dim, dtype=dtype)
    oh = oh.pin_memory() if dev.type == "cuda" else oh
    t0 = time.perf_counter()
    for step in range(a.steps):
        for layer in range(a.layers):
            q = torch.randn(1, a.heads, 1, a.head_dim, generator=g).to(dtype)
            q = q.pin_memory() if dev.type == "cuda" else q
            sess.step(q, oh, layer)
            k_new = torch.randn(1, a.kv_heads or a.heads, 1, a.head_dim, generator=g).to(dtype)
            v_new = torch.randn(1, a.kv_heads or a.heads, 1, a.head_dim, generator=g).to(dtype)
            if step % world == rank:            # round-robin owner of the new position: appends at its fill level
                sess.append_kv(layer, k_new.to(dev), v_new.to(dev))
    if dev.type == "cuda":
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if rank == 0:
        n = a.steps * a.layers
        print(f"{n} decode-attention steps over {world} rank(s) x {a.tokens_per_rank} tokens x {a.layers} layer(s), "
              f"kv={a.kv_format}: {dt / n * 1e6:.1f} us/step end to end (checksum {oh.float().sum().item():+.4f})")
    ta.cleanup()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--tokens-per-rank", type=int, default=4096)
    ap.add_argument("--heads", type=int, default=8)
    ap.add_argument("--kv-heads", type=int, default=None)
    ap.add_argument("--head-dim", type=int, default=128)
    ap.add_argument("--layers", type=int, default=2)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--kv-format", choices=["native", "fp8", "mxfp8"], default="native")
    ap.add_argument("--port", type=int, default=12361)
    a = ap.parse_args()
    if "RANK" in os.environ:
        worker(int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), a)
    elif torch.cuda.is_available() and torch.cuda.device_count() > 1:
        mp.spawn(worker, args=(torch.cuda.device_count(), a), nprocs=torch.cuda.device_count(), join=True)
    else:
        worker(0, 1, a)


if __name__ == "__main__":
    main()
