#!/usr/bin/env bash
# One gpurun call = one pass over: smoke, GPU tests, both bench arms, verbatim reference, ncu captures.
# Every step has its own timeout and the script keeps going; everything lands in gpurun_out/.
# usage: bench_tools/gpu_session.sh [tag] [steps...]   steps in: probe smoke tests bench benchN bench8pdl ref ncu_list ncu_full
#        ncu_fwd fwd bwd fwd_phases decode_bench sweep sweep_gqa tests_multi experimental bench20 benchlong sanitize configs cli
set -u
cd "$(dirname "$0")/.."
TAG=${1:-s1}; shift || true
STEPS=${*:-"probe smoke tests bench ref ncu_list ncu_full"}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
export TREE_ATTN_NO_REBUILD=1 PYTHONUNBUFFERED=1
nvidia-smi -L > "$OUT/gpus.txt" 2>&1
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,clocks.mem,power.draw,clocks_event_reasons.active --format=csv >> "$OUT/gpus.txt" 2>&1
NG=$(nvidia-smi -L | wc -l)
for st in $STEPS; do
  echo "=== $st ($(date +%T))"
  case $st in
    probe)
      timeout 900 python bench_tools/probe_all.py > "$OUT/probe.log" 2>&1; echo "rc=$?" >> "$OUT/probe.log"; cat "$OUT/probe.log" | tail -n 40;;
    smoke)
      timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1; echo "rc=$?" >> "$OUT/smoke.log"; tail -n 3 "$OUT/smoke.log";;
    tests)
      timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > "$OUT/pytest_gpu.log" 2>&1; echo "rc=$?" >> "$OUT/pytest_gpu.log"; tail -n 15 "$OUT/pytest_gpu.log";;
    bench)
      timeout 600 python bench.py --gpus 1 > "$OUT/bench_own_1.json" 2> "$OUT/bench_own_1.err"; echo "rc=$?"; tail -n 2 "$OUT/bench_own_1.json"; tail -n 3 "$OUT/bench_own_1.err"
      timeout 600 python bench.py --gpus 1 --impl reference > "$OUT/bench_ref_1.json" 2> "$OUT/bench_ref_1.err"; echo "rc=$?"; tail -n 2 "$OUT/bench_ref_1.json"; tail -n 3 "$OUT/bench_ref_1.err";;
    benchN)
      for n in 2 4 8; do
        [ "$n" -le "$NG" ] || continue
        timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n > "$OUT/bench_own_$n.json" 2> "$OUT/bench_own_$n.err"; echo "own n=$n rc=$?"; tail -n 1 "$OUT/bench_own_$n.json"; tail -n 3 "$OUT/bench_own_$n.err"
        timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700+n)) bench.py --gpus $n --impl reference --steps 20 > "$OUT/bench_ref_$n.json" 2> "$OUT/bench_ref_$n.err"; echo "ref n=$n rc=$?"; tail -n 1 "$OUT/bench_ref_$n.json"
      done;;
    ref)
      mkdir -p "$OUT/ref_verbatim"
      ( REPO=$PWD; cd "$OUT/ref_verbatim" && CUDA_VISIBLE_DEVICES=0 timeout 180 python3 "$REPO/baseline/_ref/model.py" > run_1gpu.log 2>&1; echo "rc=$?" >> run_1gpu.log
        if [ "$NG" -ge 2 ]; then CUDA_VISIBLE_DEVICES=0,1 timeout 180 python3 "$REPO/baseline/_ref/model.py" > run_2gpu.log 2>&1; echo "rc=$?" >> run_2gpu.log; fi )
      tail -n 3 "$OUT"/ref_verbatim/*.log;;
    ncu_list)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file "$OUT/launches.csv" \
        python bench.py --gpus 1 --steps 5 --warmup 3 --no-extras --no-graph > "$OUT/ncu_list.log" 2>&1; echo "rc=$?"; tail -n 5 "$OUT/launches.csv";;
    ncu_full)
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_simt -s 4 -c 1 -f -o "$OUT/prof_decode" \
        python bench.py --gpus 1 --steps 3 --warmup 3 --no-extras --no-graph > "$OUT/ncu_full.log" 2>&1; echo "rc=$?"; ls -la "$OUT" | grep ncu-rep;;
    ncu_fwd)
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 2 -c 1 -f -o "$OUT/prof_fwd" \
        python bench_tools/bench_fwd.py --seq 16384 --steps 2 --warmup 1 > "$OUT/ncu_fwd.log" 2>&1; echo "rc=$?"; ls -la "$OUT" | grep ncu-rep;;
    fwd)
      timeout 900 python bench_tools/bench_fwd.py > "$OUT/bench_fwd.log" 2>&1; echo "rc=$?"; tail -n 30 "$OUT/bench_fwd.log";;
    bench8pdl)
      timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29833 bench.py --gpus $NG --pdl --no-extras > "$OUT/bench_own_${NG}_pdl.json" 2> "$OUT/bench_own_${NG}_pdl.err"; echo "pdl n=$NG rc=$?"; tail -n 1 "$OUT/bench_own_${NG}_pdl.json" | cut -c1-400;;
    sweep_gqa)
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29812 bench_tools/sweep.py --kv-heads 8 --modes decode --seqs 131072 1048576 --out "$OUT/sweep_gqa_$NG.jsonl" > "$OUT/sweep_gqa_$NG.log" 2>&1; echo "rc=$?"; tail -n 12 "$OUT/sweep_gqa_$NG.log" | cut -c1-300;;
    sweep)
      timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29811 bench_tools/sweep.py --out "$OUT/sweep_$NG.jsonl" > "$OUT/sweep_$NG.log" 2>&1; echo "rc=$?"; tail -n 20 "$OUT/sweep_$NG.log";;
    tests_multi)
      timeout 1500 python -m pytest tests/test_gpu_multi.py -q --timeout 600 -p no:cacheprovider > "$OUT/pytest_multi.log" 2>&1; echo "rc=$?" >> "$OUT/pytest_multi.log"; tail -n 15 "$OUT/pytest_multi.log";;
    experimental)   # the cta_group::2 probe (the 2-CTA forward built on it was measured slower and removed)
      TREE_ATTN_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_probe.py -q -k 2cta -p no:cacheprovider > "$OUT/pytest_2cta_probe.log" 2>&1; echo "rc=$?" >> "$OUT/pytest_2cta_probe.log"; tail -n 8 "$OUT/pytest_2cta_probe.log";;
    decode_bench)
      timeout 300 python bench_tools/bench_decode.py --seq 131072 262144 --steps 100 > "$OUT/bench_decode_mha.log" 2>&1; tail -n 2 "$OUT/bench_decode_mha.log" | cut -c1-600
      timeout 300 python bench_tools/bench_decode.py --seq 1048576 --kv-heads 8 --steps 50 > "$OUT/bench_decode_gqa.log" 2>&1; tail -n 1 "$OUT/bench_decode_gqa.log" | cut -c1-600;;
    bwd)
      timeout 600 python bench_tools/bench_bwd.py --seq 16384 > "$OUT/bench_bwd.log" 2>&1; echo "rc=$?"; tail -n 4 "$OUT/bench_bwd.log";;
    fwd_phases)
      timeout 300 python bench_tools/prof_fwd_phases.py > "$OUT/fwd_phases.log" 2>&1; echo "rc=$?"; tail -n 4 "$OUT/fwd_phases.log";;
    scale)     # what the driver does at round end: N = 1, 2, 4, 8 back to back on ONE box, 20 timed steps, 5 warm-ups
      for n in 1 2 4 8; do
        [ "$n" -le "$NG" ] || continue
        if [ "$n" -eq 1 ]; then
          timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-extras > "$OUT/scale_own_1.json" 2> "$OUT/scale_own_1.err"
        else
          timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29850+n)) bench.py --gpus $n --steps 20 --warmup 5 --no-extras > "$OUT/scale_own_$n.json" 2> "$OUT/scale_own_$n.err"
        fi
        echo "scale n=$n rc=$?"; tail -n 1 "$OUT/scale_own_$n.json" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['ms_per_step'], d['e2e']['ms_per_step'], d['clocks'])"
      done;;
    bench1)    # same-box 1-GPU reference point for the scaling efficiency (boxes differ by several percent)
      CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-extras > "$OUT/bench20_own_1_samebox.json" 2> "$OUT/bench20_own_1_samebox.err"; echo "own20 n=1 rc=$?"; tail -n 1 "$OUT/bench20_own_1_samebox.json" | cut -c1-500;;
    bench20)   # the driver's own invocation: 20 timed steps, 5 warm-ups, both arms
      if [ "$NG" -ge 2 ]; then
        timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29841 bench.py --gpus $NG --steps 20 --warmup 5 > "$OUT/bench20_own_$NG.json" 2> "$OUT/bench20_own_$NG.err"; echo "own20 n=$NG rc=$?"; tail -n 1 "$OUT/bench20_own_$NG.json" | cut -c1-1500; tail -n 3 "$OUT/bench20_own_$NG.err" | cut -c1-300
        timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29842 bench.py --gpus $NG --steps 20 --warmup 5 --impl reference > "$OUT/bench20_ref_$NG.json" 2> "$OUT/bench20_ref_$NG.err"; echo "ref20 n=$NG rc=$?"; tail -n 1 "$OUT/bench20_ref_$NG.json" | cut -c1-400
      else
        timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > "$OUT/bench20_own_1.json" 2> "$OUT/bench20_own_1.err"; echo "own20 rc=$?"; tail -n 1 "$OUT/bench20_own_1.json" | cut -c1-1500; tail -n 3 "$OUT/bench20_own_1.err" | cut -c1-300
        timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --impl reference > "$OUT/bench20_ref_1.json" 2> "$OUT/bench20_ref_1.err"; echo "ref20 rc=$?"; tail -n 1 "$OUT/bench20_ref_1.json" | cut -c1-600
      fi;;
    benchlong)  # same config, 2000 timed steps: the 20-step value must agree with this one
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29843 bench.py --gpus $NG --steps 2000 --warmup 50 --no-extras > "$OUT/benchlong_own_$NG.json" 2> "$OUT/benchlong_own_$NG.err"; echo "long n=$NG rc=$?"; tail -n 1 "$OUT/benchlong_own_$NG.json" | cut -c1-700;;
    sanitize)
      bash bench_tools/sanitize.sh ${SAN_TOOLS:-memcheck synccheck} > "$OUT/sanitize.log" 2>&1; mkdir -p "$OUT/sanitize"; cp gpurun_out/sanitize/* "$OUT/sanitize/" 2>/dev/null; cat gpurun_out/sanitize/summary.txt;;
    configs)
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29844 bench_tools/configs.py --scale-down ${CONFIGS_SCALE:-1} --out "$OUT/configs_$NG.json" > "$OUT/configs_$NG.log" 2>&1; echo "rc=$?"; tail -n 3 "$OUT/configs_$NG.log" | cut -c1-2500;;
    cli)   # the reference UX on hardware: zero-arg spawn over every visible GPU, then torchrun --json
      timeout 300 python3 model.py > "$OUT/cli_zero_arg_$NG.log" 2>&1; echo "model.py rc=$?" | tee -a "$OUT/cli_zero_arg_$NG.log"; tail -n 6 "$OUT/cli_zero_arg_$NG.log" | cut -c1-300
      timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29845 model.py --json > "$OUT/cli_torchrun_$NG.log" 2>&1; echo "torchrun model.py rc=$?" | tee -a "$OUT/cli_torchrun_$NG.log"; tail -n 3 "$OUT/cli_torchrun_$NG.log" | cut -c1-600;;
    *) echo "unknown step $st";;
  esac
done
echo "=== done ($(date +%T))"
