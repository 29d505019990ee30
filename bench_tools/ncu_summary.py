"""Condense .ncu-rep files (ncu --set full) into the text summaries kept under profiles/:
    python bench_tools/ncu_summary.py gpurun_out/ncu/*.ncu-rep --out profiles/r2_ncu
Per report: kernel name, duration, DRAM / tensor / issue metrics, registers, shared memory, top stall reasons."""
import argparse, csv, io, os, re, subprocess, sys

KEEP = re.compile(r"^(gpu__time_duration\.sum|dram__bytes_(read|write)\.sum(\.per_second)?|gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|"
                  r"sm__pipe_tensor_cycles_active\.avg\.pct_of_peak_sustained_(active|elapsed)|sm__throughput\.avg\.pct_of_peak_sustained_elapsed|"
                  r"smsp__inst_executed\.sum|smsp__issue_active\.avg\.pct_of_peak_sustained_active|sm__warps_active\.avg\.pct_of_peak_sustained_active|"
                  r"sm__inst_executed_pipe_(fma|alu|xu|lsu|uniform|tensor.*)\.avg\.pct_of_peak_sustained_active|"
                  r"launch__(registers_per_thread|grid_size|block_size|shared_mem_per_block_dynamic|cluster.*|occupancy_limit.*)|"
                  r"sm__cycles_elapsed\.avg\.per_second|lts__t_sector_hit_rate\.pct|l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum|"
                  r"smsp__inst_executed_op_local_(ld|st)\.sum|smsp__average_warp.*_per_issue_active.*|"
                  r"smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio|sm__inst_executed_pipe_tensor.*)$")

def main():
    ap = argparse.ArgumentParser(); ap.add_argument("reps", nargs="+"); ap.add_argument("--out", required=True); ap.add_argument("--note", default="")
    a = ap.parse_args(); os.makedirs(a.out, exist_ok=True)
    for rep in a.reps:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        if len(rows) < 3: print("skip", rep); continue
        hdr, units, vals = rows[0], rows[1], rows[-1]
        name = os.path.splitext(os.path.basename(rep))[0]
        lines = [f"ncu --set full --clock-control none --import-source on, report {name}.ncu-rep (round 2, 1x B200; bench_tools/ncu_all.sh)", a.note, ""]
        for h, u, v in zip(hdr, units, vals):
            if h == "Kernel Name": lines.append(f"Kernel Name = {v}")
        stalls = []
        for h, u, v in zip(hdr, units, vals):
            if KEEP.match(h):
                if "issue_stalled" in h:
                    try: stalls.append((float(v), h))
                    except ValueError: pass
                else: lines.append(f"{h} [{u}] = {v}")
        stalls.sort(reverse=True)
        lines.append(""); lines.append("top warp stall reasons (warps stalled per issue-active cycle):")
        for v, h in stalls[:8]: lines.append(f"  {h} = {v:.3f}")
        open(os.path.join(a.out, f"ncu_{name}.txt"), "w").write("\n".join(lines) + "\n")
        print("wrote", name)

main()
