"""SASS / resource evidence of the built extension (runs on a CPU box: cuobjdump only needs the .so).

    python bench_tools/sass_report.py [--out profiles/sass]

Writes, for one representative instantiation of every kernel, its `cuobjdump -sass` listing, and a README with
  * the count of the Blackwell-specific mnemonics per kernel (UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG /
    UTMASTG = TMA, SYNCS = mbarrier, UBLKCP = cp.async.bulk, .SYS loads / stores = peer traffic inside the kernel,
    FFMA2 = packed fma.rn.f32x2);
  * registers / stack / static shared memory of EVERY instantiation (`cuobjdump -res-usage`).
"""
import argparse
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "tree_attention_b200", "csrc", "build", "_C.so")

# listing name -> regex on the demangled kernel name (first match wins)
PICK = [
    ("decode_simt_d128_r1_bf16", r"decode_simt_kernel<128, 1, 1, 0>"),       # <D, rows, bf16, fp8 KV>
    ("decode_simt_d128_r1_mxfp8", r"decode_simt_kernel<128, 1, 1, 1>"),
    ("decode_tc_d128_bf16", r"decode_tc_kernel<128, 1, 0>"),                   # <D, bf16, fp8 KV>
    ("decode_tc_d128_fp8", r"decode_tc_kernel<128, 1, 1>"),
    ("decode_swap_bf16", r"decode_swap_kernel<1, 0, 0>"),                      # <bf16, fp8 KV, block-scaled>
    ("decode_swap_fp8_per_channel", r"decode_swap_kernel<1, 1, 0>"),
    ("decode_swap_mx_block_scaled", r"decode_swap_kernel<1, 1, 1>"),
    ("attn_fwd_d128_bf16", r"attn_fwd_kernel<128, 1, 0>"),                     # <D, bf16, fused combine>
    ("attn_fwd_d128_bf16_fused_comm", r"attn_fwd_kernel<128, 1, 1>"),
    ("bwd_dq_d128_bf16", r"bwd_dq_kernel<128, 1>"),
    ("bwd_dkv_d128_bf16", r"bwd_dkv_kernel<128, 1>"),
    ("combine_oneshot", r"combine_oneshot_kernel"),
    ("combine_butterfly", r"combine_butterfly_kernel"),
    ("symm_allreduce", r"symm_allreduce_kernel"),
    ("mxfp8_seq_append", r"mxfp8_seq_append_kernel<0>"),
    ("umma_probe", r"umma_probe_kernel"),
    ("umma_bs_probe", r"umma_bs_probe_kernel"),
    ("umma_2cta_probe", r"umma_2cta_probe_kernel"),
]
NOTABLE = ("UTCHMMA", "UTCQMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMACCTL", "UBLKCP", "SYNCS",
           "FFMA2", "MUFU", "ELECT", "R2UR", "ATOMG", "REDUX", "CREDUX", "MEMBAR", "FENCE", "ERRBAR", "LDS", "STS", "LDG", "STG", "LD", "ST")


def sh(*cmd):
    return subprocess.run(cmd, check=True, capture_output=True, text=True).stdout


def demangle(names):
    out = subprocess.run(["cu++filt"], input="\n".join(names), capture_output=True, text=True, check=True).stdout.splitlines()
    out = [re.sub(r"\((?:int|bool)\)", "", o) for o in out]      # "<(int)128, (bool)1>" -> "<128, 1>"
    return dict(zip(names, out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "sass"))
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    sass = sh("cuobjdump", "-sass", SO)
    blocks = {}
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            blocks[cur] = []
        if cur is not None:
            blocks[cur].append(line)
    dm = demangle(list(blocks))
    short = lambda n: re.sub(r"\(anonymous namespace\)::|<unnamed>::|ta::|void ", "", dm[n]).split("(CUtensorMap")[0].split("(CombineParams")[0].split("(ReduceParams")[0].split("(PrepParams")[0].split("(const ")[0].split("(long long")[0]

    res = sh("cuobjdump", "-res-usage", SO)
    usage = {}
    fn = None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            fn = m.group(1)
            continue
        if fn and "REG:" in line:
            usage[fn] = dict(kv.split(":") for kv in line.split() if ":" in kv and not kv.startswith("CONSTANT"))
            fn = None

    lines = [
        "SASS listings and resource usage of tree_attention_b200/csrc/build/_C.so (sm_100a, nvcc 12.9), written by",
        "bench_tools/sass_report.py.  Blackwell-native evidence: UTC*MMA = tcgen05.mma (UTCHMMA f16/bf16, UTCQMMA fp8; the block-scaled",
        "form is the UTCQMMA with an extra tmem[] scale-factor operand; .2CTA = cta_group::2), LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG =",
        "TMA, UBLKCP = cp.async.bulk, SYNCS = mbarrier, FFMA2 = packed fma.rn.f32x2; peer traffic of the fused kernels = LD/ST with the",
        ".SYS scope on symmetric-memory pointers in the same kernel as the attention tiles.",
        "",
    ]
    for fname, pat in PICK:
        hit = next((n for n in blocks if re.search(pat, dm[n])), None)
        if hit is None:
            lines.append(f"{fname}: (no instantiation matches {pat!r})")
            continue
        body = blocks[hit]
        with open(os.path.join(a.out, fname + ".sass"), "w") as f:
            f.write("\n".join(body) + "\n")
        ops = collections.Counter()
        n_instr = 0
        sys_scope = two_cta = 0
        for ln in body:
            m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Za-z0-9_]+)*)", ln)
            if not m:
                continue
            n_instr += 1
            ops[m.group(1)] += 1
            if ".SYS" in m.group(2) and m.group(1) in ("LD", "ST", "LDG", "STG", "ATOMG", "MEMBAR", "FENCE", "ERRBAR", "CCTL"):
                sys_scope += 1
            if ".2CTA" in m.group(2):
                two_cta += 1
        cnt = ", ".join(f"{k}={ops[k]}" for k in NOTABLE if ops.get(k))
        u = usage.get(hit, {})
        lines.append(f"{fname}  [{short(hit)}]: {n_instr} instr; regs {u.get('REG', '?')}, stack {u.get('STACK', '?')} B, "
                     f"local {u.get('LOCAL', '?')} B; .SYS-scope memory ops {sys_scope}" + (f"; .2CTA ops {two_cta}" if two_cta else "") + f"; {cnt}")
    lines += ["", "Resource usage of every kernel instantiation (cuobjdump -res-usage; dynamic shared memory is set at launch):", ""]
    for n in sorted(usage, key=lambda x: dm.get(x, x)):
        if n not in dm:
            continue
        u = usage[n]
        lines.append(f"  {short(n)}: REG {u.get('REG')} STACK {u.get('STACK')} LOCAL {u.get('LOCAL')} SHARED(static) {u.get('SHARED')}")
    with open(os.path.join(a.out, "README.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines[6:6 + len(PICK)]))


if __name__ == "__main__":
    sys.exit(main())
