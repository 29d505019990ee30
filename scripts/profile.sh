#!/usr/bin/env bash
# One ncu capture of the hot kernel of a config (see /profiles/README.md for what is kept).
#   scripts/profile.sh decode|fwd|bwd [out-dir]
set -euo pipefail
cd "$(dirname "$0")/.."
what=${1:-decode}; out=${2:-gpurun_out/profile}; mkdir -p "$out"
case "$what" in
  decode) regex=decode_simt; cmd="python bench.py --steps 3 --warmup 1 --no-extras --no-graph --no-pdl" ;;
  fwd)    regex=attn_fwd;    cmd="python bench_tools/bench_fwd.py --seq 16384 --steps 1 --warmup 1 --libs 0" ;;
  bwd)    regex=bwd_d;       cmd="python bench_tools/bench_bwd.py --seq 16384" ;;
  *) echo "unknown target $what"; exit 2 ;;
esac
ncu --set full --clock-control none --import-source on -k regex:$regex -s 1 -c 1 -f -o "$out/$what" $cmd
ncu -i "$out/$what.ncu-rep" --page raw --csv > "$out/$what.raw.csv"
