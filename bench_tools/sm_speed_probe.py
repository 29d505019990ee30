"""Is the per-SM streaming rate of the decode kernel systematic (same SMs slow in every launch) or random?

    python bench_tools/sm_speed_probe.py --seq 131072 --launches 6 > gpurun_out/sm_speed.json

Traces several isolated launches of csrc/decode_simt.cu (stamps: decode_set_trace) over alternating K/V buffers and
prints, per launch, the microseconds per tile of every CTA keyed by the SM it ran on, plus the launch-to-launch
correlation of those rates.  A high correlation means a calibrated (per-SM weighted) split would remove the straggler
tail; a low one means only run-time work stealing can.
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
    ap.add_argument("--seq", type=int, nargs="*", default=[131072])
    ap.add_argument("--heads", type=int, default=32)
    ap.add_argument("--launches", type=int, default=6)
    a = ap.parse_args()
    C = _build.load()
    dev = torch.device("cuda")
    nsm = C.num_sms()
    for S in a.seq:
        g = torch.Generator(device=dev).manual_seed(0)
        q = torch.randn(1, a.heads, 1, 128, device=dev, generator=g).bfloat16()
        kvs = [(torch.randn(1, a.heads, S, 128, device=dev, generator=g).bfloat16(),
                torch.randn(1, a.heads, S, 128, device=dev, generator=g).bfloat16()) for _ in range(3)]
        out = torch.empty_like(q)
        for i in range(5):
            L.decode_attention(q, *kvs[i % 3], 0.088, out=out, return_lse=False)
        torch.cuda.synchronize()
        rates, smids, streams = [], [], []
        for i in range(a.launches):
            tr = torch.zeros(nsm * 16, dtype=torch.int64, device=dev)
            C.decode_set_trace(tr)
            L.decode_attention(q, *kvs[i % 3], 0.088, out=out, return_lse=False)
            torch.cuda.synchronize()
            C.decode_set_trace(None)
            t = tr.view(nsm, 16).cpu().double()
            mhz = float(((t[:, 8] - t[:, 1]) / (t[:, 9] - t[:, 0])).median() * 1e3)
            stream_us = (t[:, 5] - t[:, 4]) / mhz
            per_tile = stream_us / (t[:, 12] - 1).clamp(min=1)
            by_sm = torch.full((nsm,), float("nan"), dtype=torch.float64)
            by_sm[t[:, 13].long()] = per_tile
            rates.append(by_sm)
            smids.append(t[:, 13].long().tolist())
            streams.append(stream_us)
        R = torch.stack(rates)
        ok = ~torch.isnan(R).any(0)
        Rc = R[:, ok]
        corr = torch.corrcoef(Rc)
        mean_sm = Rc.mean(0)
        res = {
            "seq": S, "launches": a.launches, "sms_seen": int(ok.sum()),
            "cta_to_sm_identical_across_launches": all(s == smids[0] for s in smids),
            "launch_to_launch_corr_min": float(corr[~torch.eye(len(R), dtype=torch.bool)].min()),
            "launch_to_launch_corr_mean": float(corr[~torch.eye(len(R), dtype=torch.bool)].mean()),
            "per_launch_us_per_tile(min,mean,max)": [[round(float(r.min()), 3), round(float(r.mean()), 3), round(float(r.max()), 3)] for r in Rc],
            "per_sm_mean_us_per_tile(min,mean,max)": [round(float(mean_sm.min()), 3), round(float(mean_sm.mean()), 3), round(float(mean_sm.max()), 3)],
            # what a perfect systematic calibration would leave: residual spread after dividing out the per-SM mean
            "residual_max_over_mean_after_calibration": [round(float((r / mean_sm).max() / (r / mean_sm).mean()), 4) for r in Rc],
            "max_over_mean_now": [round(float(r.max() / r.mean()), 4) for r in Rc],
            "per_sm_mean_us_per_tile_by_smid": [round(float(x), 3) for x in mean_sm.tolist()],
            "cta_to_sm_first_launch": smids[0],
        }
        print(json.dumps(res), flush=True)
        del kvs


if __name__ == "__main__":
    main()
