#!/usr/bin/env bash
# compute-sanitizer passes over the single-GPU kernels.  Run under gpurun:
#   gpurun --timeout 1200 -- bash bench_tools/sanitize.sh [tools...]      (default: memcheck synccheck racecheck)
# One process per (kernel case, tool): a tool that aborts on its first report (synccheck does) cannot hide the other
# kernels.  Small shapes; output in gpurun_out/sanitize/<case>.<tool>.log plus a one-line-per-run summary.txt.
set -u
cd "$(dirname "$0")/.."
export TREE_ATTN_NO_REBUILD=1 PYTHONUNBUFFERED=1
OUT=gpurun_out/sanitize; mkdir -p "$OUT"
TOOLS=${*:-"memcheck synccheck racecheck"}
cat > /tmp/san_driver.py <<'PY'
import sys
import torch
from tree_attention_b200.ops import flash, local as L, quant
case = sys.argv[1]
g = torch.Generator(device="cuda").manual_seed(0)
q = torch.randn(1, 4, 1, 128, device="cuda", generator=g).bfloat16()
k = torch.randn(1, 2, 700, 128, device="cuda", generator=g).bfloat16()
v = torch.randn(1, 2, 700, 128, device="cuda", generator=g).bfloat16()
q1 = torch.randn(1, 2, 1, 128, device="cuda", generator=g).bfloat16()
if case == "decode_simt":
    o, _ = L.decode_attention(q1, k, v, 0.088, False, 0, 0, impl="simt")
    o2, _ = L.decode_attention(q1, k, v, 0.088, False, 0, 0, impl="simt", kv_len=torch.tensor([130], dtype=torch.int32, device="cuda"))
    r = o.float().abs().max().item() + o2.float().abs().max().item()
elif case == "decode_simt_mx":
    kq, vq = quant.MXFP8Tensor.from_float(k), quant.MXFP8Tensor.from_float(v)
    r = L.decode_attention_mxfp8(q1, kq, vq, 0.088)[0].float().abs().max().item()
elif case == "decode_swap":
    r = L.decode_attention(q, k, v, 0.088, False, 0, 0, impl="swap")[0].float().abs().max().item()
elif case == "decode_swap_mx":
    kq, vs = quant.MXFP8Tensor.from_float(k), quant.MXFP8SeqTensor.from_float(v)
    r = L.decode_attention_mx_tc(q, kq, vs, 0.088)[0].float().abs().max().item()
elif case == "decode_tc":
    r = L.decode_attention(q, k, v, 0.088, False, 0, 0, impl="tc")[0].float().abs().max().item()
elif case in ("fwd", "bwd"):
    q2 = torch.randn(1, 4, 200, 128, device="cuda", generator=g).bfloat16()
    of, lf = flash.attention_fwd(q2, k, v, 0.088, True, 500, 0)
    r = of.float().abs().max().item()
    if case == "bwd":
        do = torch.randn_like(q2)
        dq, dk, dv = flash.attention_bwd(q2, k, v, of, lf, do, 0.088, True, 500, 0)
        r = dq.abs().max().item()
elif case == "quant":
    a, b = quant.MXFP8Tensor.from_float(k), quant.MXFP8SeqTensor.from_float(v)
    r = a.dequantize(torch.float32).abs().max().item() + b.dequantize(torch.float32).abs().max().item()
else:
    raise SystemExit(f"unknown case {case}")
torch.cuda.synchronize()
print("ok", case, r)
PY
: > "$OUT/summary.txt"
for c in decode_simt decode_simt_mx decode_swap decode_swap_mx decode_tc fwd bwd quant; do
  for tool in $TOOLS; do
    PYTHONPATH=$PWD timeout 300 compute-sanitizer --tool $tool --print-limit 5 python /tmp/san_driver.py $c > "$OUT/$c.$tool.log" 2>&1
    rc=$?
    errs=$(grep -c "^========= .*\(error\|Error\|hazard\)" "$OUT/$c.$tool.log" || true)
    summ=$(grep "ERROR SUMMARY\|RACECHECK SUMMARY" "$OUT/$c.$tool.log" | tail -n 1)
    first=$(grep -m1 "^========= \(Barrier\|Invalid\|Race\|Error\|Warning\|ERROR:\|WARN\)" "$OUT/$c.$tool.log" | cut -c1-160)
    echo "$c $tool rc=$rc | $summ | $first" | tee -a "$OUT/summary.txt"
  done
done
