#!/usr/bin/env python3
"""Per-source-line instruction share of a kernel from an .ncu-rep (needs -lineinfo and
--import-source on).  usage: ncu_lines.py report.ncu-rep [top_n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if "Instructions Executed" in r][0]
hdr = rows[hi]
ie, smp = hdr.index("Instructions Executed"), hdr.index("# Samples")
per, tot, stot = [], 0.0, 0.0
for r in rows[hi + 1:]:
    if r and r[0].isdigit():
        def f(x):
            try:
                return float(x)
            except ValueError:
                return 0.0
        if len(r) <= max(ie, smp):
            continue
        per.append((int(r[0]), r[1][:100], f(r[ie]), f(r[smp])))
        tot += per[-1][2]
        stot += per[-1][3]
print(f"total warp-instructions {tot:.0f}, stall samples {stot:.0f}")
for ln, src, v, s in sorted(per, key=lambda p: -p[2])[:top]:
    print(f"{ln:5d} {100 * v / tot:5.1f}% inst {100 * s / max(stot, 1):5.1f}% smp  {src}")

# ---- per-phase aggregation: phases are delimited by "// ---- Pn" comments in the kernel source
import re
src_file = None
for cand in ("dump1090_rs_b200/csrc/scan7.cuh",):
    try:
        src_file = open(cand).read().splitlines()
    except OSError:
        pass
if src_file:
    marks = [(i + 1, m.group(1)) for i, l in enumerate(src_file) for m in [re.search(r"// ---- (P\w+)", l)] if m]
    k0 = [i + 1 for i, l in enumerate(src_file) if "scan7_kernel(const Scan7Params P)" in l]
    k1 = [len(src_file) + 1]
    agg = {}
    for ln, src, v, s in per:
        if k0 and k1 and k0[0] <= ln < k1[0]:
            ph = "prologue"
            for m_ln, name in marks:
                if ln >= m_ln:
                    ph = name
        else:
            ph = "helpers:" + ("mag" if "mag" in src or ln < 170 and ln > 140 else "other")
            # helper functions are attributed by content below
            txt = src
            if any(t in txt for t in ("kTab56", "t[f &", "mulx", "s <<= 1", "0x1000000u", "syn", "K_PAR", "msg_bits", "kNoneMarker", "df ==", "bit &")):
                ph = "helpers:crc/classify (P4)"
            elif any(t in txt for t in ("fmaf", "fmul", "byte_perm", "rsqrt", "fadd_rz", "0x80008000", "fminf")):
                ph = "helpers:mag (P1)"
            elif any(t in txt for t in ("pp[", "cs ==", "high", "noise", "sig ", "mx ", "surv[")):
                ph = "helpers:gate (P3)"
            elif any(t in txt for t in ("r2", "funnelshift_r(p[0]")):
                ph = "helpers:plane_term (P3)"
            elif any(t in txt for t in ("atomic", "ev_", "hash32")):
                ph = "helpers:events (P4)"
            else:
                ph = "helpers:other"
        a = agg.setdefault(ph, [0.0, 0.0])
        a[0] += v
        a[1] += s
    print("\nper phase:")
    for ph, (v, s) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"  {ph:32s} {100 * v / tot:5.1f}% inst  {100 * s / max(stot, 1):5.1f}% stall samples")
