"""Run the tcgen05 layout probes and print an error matrix (one process per group so that a faulting
configuration cannot take the others down).  usage: probe_all.py [group]"""
import subprocess
import sys

import torch

GROUPS = {
    "ss_k": [(n, k, False, False) for n, k in [(128, 64), (128, 128), (64, 128), (256, 128), (128, 256), (16, 128), (32, 64)]],
    "ss_mn": [(n, k, True, False) for n, k in [(128, 64), (128, 128), (64, 128), (256, 128), (128, 256)]],
    "ts_k": [(n, k, False, True) for n, k in [(128, 64), (128, 128), (64, 128), (16, 128)]],
    "ts_mn": [(n, k, True, True) for n, k in [(128, 64), (128, 128), (64, 128), (256, 128), (128, 256)]],
}


def run_group(name):
    from tree_attention_b200 import _build

    C = _build.load()
    for (n, k, mn, ts) in GROUPS[name]:
        g = torch.Generator(device="cuda").manual_seed(n * 1000 + k)
        a = torch.randn(128, k, device="cuda", generator=g).bfloat16()
        b = torch.randn((k, n) if mn else (n, k), device="cuda", generator=g).bfloat16()
        c = torch.full((128, n), float("nan"), device="cuda", dtype=torch.float32)
        try:
            C.umma_probe(a, b, c, mn, ts)
            torch.cuda.synchronize()
            exp = a.float() @ (b.float() if mn else b.float().t())
            err = (c - exp).abs().max().item()
            rel = err / exp.abs().max().item()
            print(f"{name} N={n} K={k} b_mn={int(mn)} a_tmem={int(ts)}: max_abs_err={err:.4g} rel={rel:.3g} "
                  f"{'OK' if rel < 5e-3 else 'WRONG'}", flush=True)
            if rel >= 5e-3:
                # diagnose: which rows / cols match
                ok_rows = ((c - exp).abs().max(dim=1).values < 1e-2 * exp.abs().max()).sum().item()
                ok_cols = ((c - exp).abs().max(dim=0).values < 1e-2 * exp.abs().max()).sum().item()
                print(f"    rows ok: {ok_rows}/128, cols ok: {ok_cols}/{n}, nan: {torch.isnan(c).sum().item()}", flush=True)
        except Exception as e:
            print(f"{name} N={n} K={k} b_mn={int(mn)} a_tmem={int(ts)}: EXCEPTION {type(e).__name__}: {str(e)[:200]}", flush=True)
            return 1
    return 0


if __name__ == "__main__":
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    if len(sys.argv) > 1:
        sys.exit(run_group(sys.argv[1]))
    for g in GROUPS:
        r = subprocess.run([sys.executable, __file__, g], timeout=240)
        print(f"group {g}: rc={r.returncode}", flush=True)
