"""Context-parallel training step: the context tokens are sharded over ranks, attention runs as tree attention
(forward: per-rank partial + tree combine; backward: dK/dV local, dQ summed over ranks).

    python examples/train_context_parallel.py --world 2           # CPU: gloo, two processes
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 examples/train_context_parallel.py

The toy model is one ``TreeSelfAttention`` block + a linear head.  Query tokens are replicated; every rank projects its
own slice of the context to K/V.  Parameters are replicated, so their gradients are averaged with one all-reduce per
step (what DDP would do).  The script checks the sharded loss and gradients against a single-process run over the full
context.

``--causal``: causal SELF-attention over the whole sequence (queries = all tokens, replicated) with the K/V tokens in
ZIGZAG shards (rank r projects chunks r and 2W-1-r, ``ta.zigzag_shard``): every rank then does the same share of the
causal work instead of rank 0 doing twice the average."""
from __future__ import annotations

import argparse
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tree_attention_b200 as ta  # noqa: E402
from tree_attention_b200.models.tree_attention import TreeSelfAttention  # noqa: E402


class Block(nn.Module):
    def __init__(self, e: int, h: int, hkv: int, dtype, device, causal: bool = False):
        super().__init__()
        self.attn = TreeSelfAttention(e, h, hkv, causal=causal, dtype=dtype, device=device,
                                      kv_layout="zigzag" if causal else "contiguous")
        self.head = nn.Linear(e, 1, dtype=dtype, device=device)

    def forward(self, x_q, x_kv, kv_offset=None):
        return self.head(x_q + self.attn(x_q, x_kv, kv_offset=kv_offset)).float().square().mean()


def worker(rank: int, world: int, a) -> None:
    dev = torch.device(f"cuda:{rank}" if torch.cuda.is_available() else "cpu")
    ta.setup(rank, world, master_port=a.port)
    dtype = torch.bfloat16 if dev.type == "cuda" else torch.float32
    torch.manual_seed(0)                        # identical parameters and data on every rank
    model = Block(a.embed, a.heads, a.kv_heads, dtype, dev, causal=a.causal)
    x_q = torch.randn(1, a.q_tokens, a.embed, device=dev).to(dtype)
    ctx = torch.randn(1, a.ctx_tokens * world, a.embed, device=dev).to(dtype)
    x_kv = ctx[:, rank * a.ctx_tokens:(rank + 1) * a.ctx_tokens]      # this rank's slice of the context
    if a.causal:                                                       # self-attention: every token is a query
        x_q = ctx
        x_kv = ta.zigzag_shard(ctx, rank, world, dim=1)                # chunks r and 2W-1-r of the token sequence
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    for step in range(a.steps):
        opt.zero_grad(set_to_none=True)
        loss = model(x_q, x_kv, kv_offset=None if a.causal else rank * a.ctx_tokens)
        loss.backward()
        if world > 1:
            for p in model.parameters():
                # q_proj / o_proj / head see the same (replicated) activations on every rank: their grads are already
                # identical; kv_proj sees a different context slice per rank: its grads SUM over ranks
                is_kv = p is model.attn.kv_proj.weight or p is getattr(model.attn.kv_proj, "bias", None)
                dist.all_reduce(p.grad, op=dist.ReduceOp.SUM)
                if not is_kv:
                    p.grad /= world
        if step == 0 and a.check:
            ref = Block(a.embed, a.heads, a.kv_heads, dtype, dev, causal=a.causal)
            ref.load_state_dict(model.state_dict())
            saved = ta.get_runtime().world_size
            # single-process reference over the full context: plain attention through the same module on a 1-rank view
            nq = x_q.shape[1]
            q = ref.attn.q_proj(x_q).view(1, nq, a.heads, -1).transpose(1, 2)
            kv = ref.attn.kv_proj(ctx).view(1, ctx.shape[1], 2, a.kv_heads, q.shape[-1])
            k, v = kv[:, :, 0].transpose(1, 2), kv[:, :, 1].transpose(1, 2)
            g = a.heads // a.kv_heads
            kf, vf = k.float().repeat_interleave(g, dim=1), v.float().repeat_interleave(g, dim=1)
            logits = q.float() @ kf.transpose(-1, -2) * q.shape[-1] ** -0.5
            if a.causal:
                logits = logits.masked_fill(torch.ones(nq, nq, dtype=torch.bool, device=dev).triu(1), float("-inf"))
            p_attn = torch.softmax(logits, dim=-1)
            o = (p_attn @ vf).to(dtype)
            y = ref.attn.o_proj(o.transpose(1, 2).reshape(1, nq, -1))
            loss_ref = ref.head(x_q + y).float().square().mean()
            loss_ref.backward()
            tol = 5e-2 if dtype == torch.bfloat16 else 1e-4
            assert abs(loss.item() - loss_ref.item()) < tol * max(1.0, abs(loss_ref.item())), (loss.item(), loss_ref.item())
            for (n, p), (_, pr) in zip(model.named_parameters(), ref.named_parameters()):
                err = (p.grad.float() - pr.grad.float()).abs().max().item()
                assert err < tol * max(1.0, pr.grad.float().abs().max().item()), (n, err)
            if rank == 0:
                print(f"step 0: sharded loss {loss.item():.6f} == full-context loss {loss_ref.item():.6f}; gradients match "
                      f"({saved} rank(s))")
        opt.step()
        if rank == 0:
            print(f"step {step}: loss {loss.item():.6f}")
    ta.cleanup()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, default=None, help="processes to spawn (default: one per GPU, or 1 on CPU)")
    ap.add_argument("--embed", type=int, default=256)
    ap.add_argument("--heads", type=int, default=2)
    ap.add_argument("--kv-heads", type=int, default=1)
    ap.add_argument("--q-tokens", type=int, default=128)
    ap.add_argument("--ctx-tokens", type=int, default=256, help="context tokens PER RANK")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--check", action=argparse.BooleanOptionalAction, default=True)
    ap.add_argument("--causal", action="store_true", help="causal self-attention over the whole sequence, zigzag K/V shards")
    ap.add_argument("--port", type=int, default=12362)
    a = ap.parse_args()
    if "RANK" in os.environ:
        worker(int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), a)
        return
    world = a.world or (torch.cuda.device_count() if torch.cuda.is_available() else 1)
    if world > 1:
        mp.spawn(worker, args=(world, a), nprocs=world, join=True)
    else:
        worker(0, 1, a)


if __name__ == "__main__":
    main()
