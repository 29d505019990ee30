"""Per-CTA timeline of the streaming decode kernel (csrc/decode_simt.cu trace stamps): where does a step's time go?

    python bench_tools/trace_decode.py --seq 16384 131072 [--heads 32]

For each shard length: times back-to-back launches (CUDA events), then traces one launch and prints, over the CTAs,
min / median / max of every interval (SM cycles converted with the sampled clock, plus globaltimer for the whole kernel).
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tree_attention_b200 import _build
from tree_attention_b200.ops import local as L


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seq", type=int, nargs="*", default=[16384, 32768, 65536, 131072])
    ap.add_argument("--heads", type=int, default=32)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--steps", type=int, default=50)
    a = ap.parse_args()
    C = _build.load()
    dev = torch.device("cuda")
    nsm = C.num_sms()
    for S in a.seq:
        g = torch.Generator(device=dev).manual_seed(0)
        q = torch.randn(1, a.heads, 1, a.dim, device=dev, generator=g).bfloat16()
        nbuf = max(2, -(-4 * (126 << 20) // (2 * a.heads * S * a.dim * 2)))
        kvs = [(torch.randn(1, a.heads, S, a.dim, device=dev, generator=g).bfloat16(),
                torch.randn(1, a.heads, S, a.dim, device=dev, generator=g).bfloat16()) for _ in range(min(nbuf, 8))]
        out = torch.empty_like(q)
        res = {"seq": S, "kv_mb": 2 * a.heads * S * a.dim * 2 / 1e6}
        for pdl in (0, 2):
            for i in range(10):
                L.decode_attention(q, *kvs[i % len(kvs)], 0.088, out=out, return_lse=False, pdl=pdl)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(a.steps):
                L.decode_attention(q, *kvs[i % len(kvs)], 0.088, out=out, return_lse=False, pdl=pdl)
            e1.record()
            torch.cuda.synchronize()
            res[f"us_per_step_pdl{pdl}"] = e0.elapsed_time(e1) / a.steps * 1e3
        # traced launches (after a warm one), isolated by synchronisation; the LAST of three is analysed, all are dumped
        tr = torch.zeros(nsm * 16, dtype=torch.int64, device=dev)
        C.decode_set_trace(tr)
        L.decode_attention(q, *kvs[0], 0.088, out=out, return_lse=False)
        torch.cuda.synchronize()
        tr.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.decode_attention(q, *kvs[1 % len(kvs)], 0.088, out=out, return_lse=False)
        e1.record()
        torch.cuda.synchronize()
        C.decode_set_trace(None)
        t = tr.view(nsm, 16).cpu().double()
        t = t[t[:, 0] > 0]
        g0 = t[:, 0].min()
        cyc = (t[:, 8] - t[:, 1])
        ns = (t[:, 9] - t[:, 0])
        mhz = float((cyc / ns).median() * 1e3)
        res["traced_kernel_us_event"] = e0.elapsed_time(e1) * 1e3
        res["sm_mhz_from_stamps"] = mhz

        def stat(x):
            x = x.sort().values
            return [round(float(v), 2) for v in (x[0], x[len(x) // 2], x[-1])]

        us = lambda c: c / mhz
        res["cta_start_skew_us(min,med,max)"] = stat((t[:, 0] - g0) / 1e3)
        res["prologue_us"] = stat(us(t[:, 2] - t[:, 1]))
        res["entry_to_first_tma_issue_us"] = stat(us(t[:, 10] - t[:, 1]))
        res["entry_to_first_tile_landed_us"] = stat(us(t[:, 4] - t[:, 1]))
        res["stream_first_tile_to_last_consumed_us"] = stat(us(t[:, 5] - t[:, 4]))
        res["last_finalize_us"] = stat(us(t[:, 6] - t[:, 5]))
        res["drain_us"] = stat(us(t[:, 8] - t[:, 6]))
        res["cta_total_us"] = stat(ns / 1e3)
        res["kernel_span_us(first entry -> last end)"] = float((t[:, 9].max() - g0) / 1e3)
        res["tiles_per_cta"] = stat(t[:, 12])
        res["producer_issue_span_us"] = stat(us(t[:, 11] - t[:, 10]))
        per_tile = us(t[:, 5] - t[:, 4]) / (t[:, 12] - 1).clamp(min=1)
        res["us_per_tile"] = stat(per_tile)
        sw = t[t[:, 14] > 0]
        if len(sw):
            fin = (sw[:, 14].long() & 0xffffffff).double()
            ldq = (sw[:, 14].long() >> 32).double()
            wt = (sw[:, 15].long() & 0xffffffff).double()
            pos = (sw[:, 15].long() >> 32).double()
            res["switch_finalize_us"] = stat(us(fin))
            res["switch_load_q_us"] = stat(us(ldq))
            res["switch_wait_next_tile_us"] = stat(us(wt))
            res["switch_tile_index"] = stat(pos)
            # per-tile rate before / after the switch for the switching CTAs
            before = us(sw[:, 7] - sw[:, 4]) / pos.clamp(min=1)
            after = us(sw[:, 5] - sw[:, 7]) / (sw[:, 12] - pos).clamp(min=1)
            res["switch_cta_us_per_tile_before"] = stat(before)
            res["switch_cta_us_per_tile_after"] = stat(after)
        if os.environ.get("TRACE_DUMP"):
            torch.save(t, os.path.join(os.environ["TRACE_DUMP"], f"trace_{S}.pt"))
        print(json.dumps(res), flush=True)
        del kvs


if __name__ == "__main__":
    main()
