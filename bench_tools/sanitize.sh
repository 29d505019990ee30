#!/usr/bin/env bash
# compute-sanitizer passes over the single-GPU kernels.  Run under gpurun:
#   gpurun --timeout 900 -- bash bench_tools/sanitize.sh
# memcheck + synccheck on small shapes (racecheck does not model TMA/tcgen05 async proxies and is reported
# separately); output in gpurun_out/sanitize/.
set -u
cd "$(dirname "$0")/.."
export TREE_ATTN_NO_REBUILD=1 PYTHONUNBUFFERED=1
OUT=gpurun_out/sanitize; mkdir -p "$OUT"
cat > /tmp/san_driver.py <<'PY'
import torch
from tree_attention_b200.ops import flash, local as L, reference as ref, quant
g = torch.Generator(device="cuda").manual_seed(0)
q = torch.randn(1, 4, 1, 128, device="cuda", generator=g).bfloat16()
k = torch.randn(1, 2, 700, 128, device="cuda", generator=g).bfloat16()
v = torch.randn(1, 2, 700, 128, device="cuda", generator=g).bfloat16()
o, l = L.decode_attention(q, k, v, 0.088, False, 0, 0)
kq, vq = quant.MXFP8Tensor.from_float(k), quant.MXFP8Tensor.from_float(v)
o8, _ = L.decode_attention_mxfp8(q, kq, vq, 0.088)
q2 = torch.randn(1, 4, 200, 128, device="cuda", generator=g).bfloat16()
of, lf = flash.attention_fwd(q2, k, v, 0.088, True, 500, 0)
do = torch.randn_like(q2)
dq, dk, dv = flash.attention_bwd(q2, k, v, of, lf, do, 0.088, True, 500, 0)
torch.cuda.synchronize()
print("ok", o.float().abs().max().item(), of.float().abs().max().item(), dq.abs().max().item())
PY
for tool in memcheck synccheck; do
  PYTHONPATH=$PWD timeout 600 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_driver.py > "$OUT/$tool.log" 2>&1
  echo "== $tool rc=$?"; tail -n 6 "$OUT/$tool.log"
done
