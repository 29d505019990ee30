"""Multi-process integration on CPU: gloo, world sizes 2/3/4 on localhost.
Oracle: monolithic fp64 attention over the concatenated KV."""
import pytest
import torch
import torch.distributed as dist

from _dist_utils import run_distributed


def _gather_kv(k, v, world):
    ks = [torch.empty_like(k) for _ in range(world)]
    vs = [torch.empty_like(v) for _ in range(world)]
    dist.all_gather(ks, k)
    dist.all_gather(vs, v)
    return torch.cat(ks, 2), torch.cat(vs, 2)


def _worker(rank, world, cfg):
    import tree_attention_b200 as ta
    from tree_attention_b200.ops import reference as ref

    b, hq, hkv, sq, s, d, causal = cfg
    q, k, v = ta.make_data((b, hq, s, d), rank, "cpu", dtype=torch.float32, sq=sq, num_kv_heads=hkv, log=False)
    kf, vf = _gather_kv(k, v, world)
    o_ref, l_ref = ref.attention_ref(q, kf, vf, causal=causal)
    for sched in ("allreduce3", "allgather", "butterfly", "oneshot", "ring"):
        out, lse = ta.tree_attention(q, k, v, causal=causal, return_lse=True, schedule=sched)
        assert torch.allclose(out.double(), o_ref, atol=1e-5), (sched, (out.double() - o_ref).abs().max())
        assert torch.allclose(lse.double(), l_ref, atol=1e-5), sched
        # every rank holds the same answer
        outs = [torch.empty_like(out) for _ in range(world)]
        dist.all_gather(outs, out.contiguous())
        for o in outs:
            assert torch.allclose(o, outs[0], atol=1e-6)
        if sched in ("allgather", "butterfly", "oneshot", "ring") and (world & (world - 1)) == 0:
            for o in outs:
                assert torch.equal(o, outs[0]), f"{sched}: not bitwise identical across ranks"
    # reference-compatible shim (scale 1.0, non-causal)
    out = ta.tree_decode(q[:, :, :1], k, v, rank, world, torch.device("cpu"))
    o_ref1, _ = ref.attention_ref(q[:, :, :1], kf, vf, softmax_scale=1.0)
    assert torch.allclose(out.double(), o_ref1, atol=1e-5)


@pytest.mark.parametrize("world", [2, 3, 4])
def test_tree_attention_gloo(world, port):
    run_distributed(_worker, world, port, args=((2, 4, 2, 3, 24, 16, True),))


def test_baseline_config_1_single_head_seq1024_world2(port):
    """BASELINE.json config #1: tree_attention() single-head seq=1024 world_size=2 on CPU/gloo."""
    run_distributed(_worker, 2, port, args=((1, 1, 1, 1, 512, 64, False),))


def _worker_unequal(rank, world):
    """Unequal shards + explicit kv_offset, including a rank whose shard is fully masked."""
    import tree_attention_b200 as ta
    from tree_attention_b200.ops import reference as ref

    lens = [5, 9, 2][:world]
    g = torch.Generator().manual_seed(0)
    q = torch.randn(1, 2, 1, 8, generator=g)
    kf = torch.randn(1, 2, sum(lens), 8, generator=g)
    vf = torch.randn(1, 2, sum(lens), 8, generator=g)
    off = sum(lens[:rank])
    k, v = kf[:, :, off:off + lens[rank]].contiguous(), vf[:, :, off:off + lens[rank]].contiguous()
    q_pos = 7  # rank 2 (positions 14..15) is fully masked, rank 1 partially
    out = ta.tree_attention(q, k, v, causal=True, kv_offset=off, q_offset=q_pos, schedule="allgather")
    o_ref, _ = ref.attention_partial_ref(q, kf, vf, causal=True, q_pos0=q_pos)
    assert torch.allclose(out, o_ref, atol=1e-5)


def test_unequal_shards_and_masked_rank(port):
    run_distributed(_worker_unequal, 3, port)


def _worker_quantised(rank, world):
    """fp8 KV caches through the distributed API on CPU (de-quantising path): every format, every rank agrees."""
    import tree_attention_b200 as ta
    from tree_attention_b200.ops import quant, reference as ref

    q, k, v = ta.make_data((1, 4, 160, 128), rank, "cpu", dtype=torch.float32, sq=2, num_kv_heads=2, log=False)
    for name, kq, vq in (
        ("channel", quant.FP8ChannelTensor.from_float(k), quant.FP8ChannelTensor.from_float(v)),
        ("mx", quant.MXFP8Tensor.from_float(k), quant.MXFP8Tensor.from_float(v)),
        ("mx_seq", quant.MXFP8Tensor.from_float(k), quant.MXFP8SeqTensor.from_float(v)),
    ):
        kf, vf = _gather_kv(kq.dequantize(), vq.dequantize(), world)
        o_ref, _ = ref.attention_ref(q, kf, vf, causal=True)
        out = ta.tree_attention(q, kq, vq, causal=True)
        assert torch.allclose(out.double(), o_ref, atol=1e-5), (name, (out.double() - o_ref).abs().max())
        outs = [torch.empty_like(out) for _ in range(world)]
        dist.all_gather(outs, out.contiguous())
        for o in outs:
            assert torch.allclose(o, outs[0], atol=1e-6), name


def test_quantised_kv_caches_gloo(port):
    run_distributed(_worker_quantised, 2, port)


def _worker_session(rank, world):
    """TreeDecodeSession on CPU ranks with FILL LEVELS: the tail of the preallocated shard holds garbage and must not
    take part in the softmax; step, append on the owner rank (which raises its level), step again -- against the oracle
    over the filled rows only (the old test zero-filled the tail and encoded the dilution)."""
    import tree_attention_b200 as ta
    from tree_attention_b200.models.decoder import TreeDecodeSession
    from tree_attention_b200.ops import reference as ref

    cap = 96
    used = 90 if rank == world - 1 else cap     # the last rank's shard is still filling up
    q, k, v = ta.make_data((1, 4, cap, 32), rank, "cpu", dtype=torch.float32, num_kv_heads=2, log=False)
    k[:, :, used:] = 1e3                        # garbage in the unwritten tail: must be ignored
    v[:, :, used:] = -1e3
    sess = TreeDecodeSession([(k, v)], softmax_scale=0.2, q_shape=(1, 4, 1, 32), backend="auto", kv_lens=[used])
    oh = torch.empty(1, 4, 1, 32)

    def oracle():
        n = sess.kv_len_host[0]
        lens = [None] * world
        dist.all_gather_object(lens, n)
        ks, vs = [None] * world, [None] * world
        dist.all_gather_object(ks, k)
        dist.all_gather_object(vs, v)
        kf = torch.cat([ks[r][:, :, : lens[r]] for r in range(world)], 2)
        vf = torch.cat([vs[r][:, :, : lens[r]] for r in range(world)], 2)
        return ref.attention_ref(q, kf, vf, softmax_scale=0.2)[0]

    out0 = sess.step(q, oh, 0).clone()
    assert torch.allclose(out0.double(), oracle(), atol=1e-5)
    g = torch.Generator().manual_seed(7)
    k_new, v_new = torch.randn(1, 2, 1, 32, generator=g) * 3, torch.randn(1, 2, 1, 32, generator=g) * 3
    if rank == world - 1:                      # the owner of the new position appends at its fill level
        sess.append_kv(0, k_new, v_new)
        assert sess.kv_len_host[0] == used + 1
    out1 = sess.step(q, oh, 0).clone()
    assert torch.allclose(out1.double(), oracle(), atol=1e-5)
    assert (out1 - out0).abs().max() > 1e-6


def _worker_kv_len(rank, world):
    """tree_attention(kv_len=...): ragged fill levels, including an EMPTY shard, on every collective schedule."""
    import tree_attention_b200 as ta
    from tree_attention_b200.ops import reference as ref

    cap = 64
    lens = [cap, 17, 0, 40][:world]
    q, k, v = ta.make_data((2, 4, cap, 32), rank, "cpu", dtype=torch.float32, num_kv_heads=2, log=False)
    k[:, :, lens[rank]:] = float("nan")        # unwritten rows hold anything
    v[:, :, lens[rank]:] = float("nan")
    ks, vs = [None] * world, [None] * world
    dist.all_gather_object(ks, k)
    dist.all_gather_object(vs, v)
    kf = torch.cat([ks[r][:, :, : lens[r]] for r in range(world)], 2)
    vf = torch.cat([vs[r][:, :, : lens[r]] for r in range(world)], 2)
    o_ref, l_ref = ref.attention_ref(q, kf, vf, softmax_scale=0.3)
    for sched in ("allgather", "allreduce3", "butterfly"):
        o, l = ta.tree_attention(q, k, v, softmax_scale=0.3, backend="gloo", schedule=sched, return_lse=True,
                                 kv_len=lens[rank])
        assert torch.allclose(o.double(), o_ref, atol=1e-5), sched
        assert torch.allclose(l.double(), l_ref, atol=1e-5), sched


def test_kv_len_ragged_gloo(port):
    run_distributed(_worker_kv_len, 4, port)


def test_decode_session_gloo(port):
    run_distributed(_worker_session, 2, port)


def _worker_sharded_output(rank, world):
    """output='sharded': this rank's block of query rows, whole 128-row tiles per rank (CPU path: slices the replicated
    result; the fused GPU path produces the same layout with a reduce-scatter inside the kernel)."""
    import tree_attention_b200 as ta
    from tree_attention_b200.ops import reference as ref

    sq = 300
    q, k, v = ta.make_data((1, 2, 64, 16), rank, "cpu", dtype=torch.float32, sq=sq, log=False)
    o_full, l_full = ta.tree_attention(q, k, v, return_lse=True, backend="gloo")
    o_sh, l_sh = ta.tree_attention(q, k, v, return_lse=True, backend="gloo", output="sharded")
    n = ((sq + 127) // 128 + world - 1) // world * 128
    assert o_sh.shape == (1, 2, n, 16) and l_sh.shape == (1, 2, n)
    lo, hi = min(rank * n, sq), min((rank + 1) * n, sq)
    assert torch.equal(o_sh[:, :, : hi - lo], o_full[:, :, lo:hi])
    assert torch.equal(l_sh[:, :, : hi - lo], l_full[:, :, lo:hi])


def test_sharded_output_gloo(port):
    run_distributed(_worker_sharded_output, 2, port)


def _worker_zigzag(rank, world):
    """kv_layout="zigzag": every rank holds chunks r and 2W-1-r of the causal sequence; the result (forward and the
    gradients) equals plain causal attention over the full sequence."""
    import torch.distributed as dist
    import tree_attention_b200 as ta
    from tree_attention_b200.ops import reference as ref
    from tree_attention_b200.ops.autograd import tree_attention_func

    g = torch.Generator().manual_seed(3)
    S, hq, hkv, d = 16 * world, 4, 2, 16
    q_full = torch.randn(1, hq, S, d, generator=g)
    k_full = torch.randn(1, hkv, S, d, generator=g)
    v_full = torch.randn(1, hkv, S, d, generator=g)
    do = torch.randn(1, hq, S, d, generator=g)
    k = ta.zigzag_shard(k_full, rank, world).clone().requires_grad_(True)
    v = ta.zigzag_shard(v_full, rank, world).clone().requires_grad_(True)
    # round trip of the layout helpers
    ks = [torch.empty_like(k) for _ in range(world)]
    dist.all_gather(ks, k.detach())
    assert torch.equal(ta.zigzag_unshard(ks), k_full)
    # oracle: causal attention over the whole sequence on one rank, with autograd
    qo, ko, vo = (t.clone().requires_grad_(True) for t in (q_full, k_full, v_full))
    o_ref, lse_ref = ref.attention_partial_ref(qo, ko, vo, causal=True, q_pos0=0)
    o_ref.backward(do)
    for sched in ("allgather", "butterfly"):
        o, lse = ta.tree_attention(q_full, k.detach(), v.detach(), causal=True, return_lse=True, kv_layout="zigzag",
                                   backend="gloo", schedule=sched)
        assert torch.allclose(o, o_ref.detach(), atol=2e-5), (sched, (o - o_ref).abs().max())
        assert torch.allclose(lse, lse_ref.detach(), atol=2e-5)
    # contiguous interpretation of the same tensors must differ (the layout really is honoured)
    o_wrong = ta.tree_attention(q_full, k.detach(), v.detach(), causal=True, backend="gloo")
    assert not torch.allclose(o_wrong, o_ref.detach(), atol=1e-3)
    # sharded output + the module front-end + decode-convention query block (last rows only)
    o_sh = ta.tree_attention(q_full, k.detach(), v.detach(), causal=True, kv_layout="zigzag", backend="gloo", output="sharded")
    n = o_sh.shape[2]
    lo, hi = min(rank * n, S), min((rank + 1) * n, S)
    assert torch.allclose(o_sh[:, :, : hi - lo], o_ref.detach()[:, :, lo:hi], atol=2e-5)
    o_tail = ta.tree_attention(q_full[:, :, -5:], k.detach(), v.detach(), causal=True, kv_layout="zigzag", backend="gloo")
    assert torch.allclose(o_tail, o_ref.detach()[:, :, -5:], atol=2e-5)
    # backward: dK / dV of my two chunks, dQ summed over ranks
    q = q_full.clone().requires_grad_(True)
    o2 = tree_attention_func(q, k, v, causal=True, backend="gloo", kv_layout="zigzag")
    o2.backward(do)
    assert torch.allclose(q.grad, qo.grad, atol=5e-5), (q.grad - qo.grad).abs().max()
    assert torch.allclose(k.grad, ta.zigzag_shard(ko.grad, rank, world), atol=5e-5)
    assert torch.allclose(v.grad, ta.zigzag_shard(vo.grad, rank, world), atol=5e-5)


@pytest.mark.parametrize("world", [2, 3])
def test_zigzag_causal_layout_gloo(world, port):
    run_distributed(_worker_zigzag, world, port)
