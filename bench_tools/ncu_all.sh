#!/usr/bin/env bash
# ncu --set full capture of the hot kernels, one .ncu-rep per kernel in gpurun_out/ncu/ (1 GPU; never under a multi-rank
# command).  usage (under gpurun): bash bench_tools/ncu_all.sh [targets...]
set -u
cd "$(dirname "$0")/.."
export TREE_ATTN_NO_REBUILD=1 PYTHONUNBUFFERED=1
OUT=gpurun_out/ncu; mkdir -p "$OUT"
declare -A KREGEX=( [decode_simt]=decode_simt_kernel [decode_simt_shard]=decode_simt_kernel [decode_swap]=decode_swap_kernel
                    [decode_swap_mx]=decode_swap_kernel [decode_tc]=decode_tc_kernel [fwd]=attn_fwd_kernel [fwd_causal]=attn_fwd_kernel
                    [bwd_dq]=bwd_dq_kernel [bwd_dkv]=bwd_dkv_kernel [quant]=quant_mxfp8 )
TARGETS=${*:-"decode_simt decode_simt_shard decode_swap decode_swap_mx decode_tc fwd bwd_dq bwd_dkv"}
for t in $TARGETS; do
  prog=$t; [ "$t" = bwd_dq ] && prog=bwd; [ "$t" = bwd_dkv ] && prog=bwd
  timeout 240 ncu --set full --clock-control none --import-source on -k regex:${KREGEX[$t]} -s 2 -c 1 -f -o "$OUT/$t" \
      python bench_tools/ncu_targets.py $prog > "$OUT/$t.log" 2>&1
  echo "$t rc=$? $(ls -la $OUT/$t.ncu-rep 2>/dev/null | awk '{print $5}') bytes"
done
