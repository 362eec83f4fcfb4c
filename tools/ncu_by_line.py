#!/usr/bin/env python
"""
Aggregate an ncu source-page CSV (SASS view) by CUDA source line.

    ncu -i prof.ncu-rep --page source --csv > prof_src.csv
    cuobjdump -xelf all libd4b200.so ; nvdisasm -g -c *.cubin > dis.txt
    python tools/ncu_by_line.py prof_src.csv dis.txt 'small_kernelIdLb0' [min_pct [source_dir]]

(source_dir: directory holding the .cuh/.cu files of the profiled build, default tad_dftd4_b200/csrc)

ncu's CSV source page carries per-SASS-instruction samples but no line numbers;
nvdisasm -g carries the line table.  Both list the kernel's instructions in
address order, so they are zipped by position (opcode text is cross-checked).
"""
import csv
import re
import sys
from collections import defaultdict


def parse_dis(path, func_pat):
    lines = open(path).read().splitlines()
    out = []  # (file, line, sass)
    infunc = False
    cur = ("?", 0)
    for ln in lines:
        if ln.startswith("//--------------------- .text."):
            infunc = re.search(func_pat, ln) is not None
            continue
        if not infunc:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            out.append((cur[0], cur[1], m.group(2).strip()))
    return out


def main():
    src_csv, dis, pat = sys.argv[1:4]
    min_pct = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
    src_dir = sys.argv[5] if len(sys.argv) > 5 else "/root/repo/tad_dftd4_b200/csrc"
    rows = list(csv.reader(open(src_csv)))
    hdr = rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    body = [r for r in rows[2:] if len(r) > ci["# Samples"]]
    sass = parse_dis(dis, pat)
    if len(sass) != len(body):
        print(f"warning: {len(sass)} disassembled vs {len(body)} profiled instructions", file=sys.stderr)
    agg = defaultdict(lambda: [0.0, 0.0, 0.0, defaultdict(float)])
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot_s = tot_i = 0.0
    for (f, line, text), r in zip(sass, body):
        s = float(r[ci["# Samples"]] or 0)
        ins = float(r[ci["Instructions Executed"]] or 0)
        op = r[ci["Source"]].split()[0] if r[ci["Source"]].split() else ""
        a = agg[(f, line)]
        a[0] += s
        a[1] += ins
        if re.match(r"@?!?P?\d*\s*D(FMA|MUL|ADD|SETP)|^D(FMA|MUL|ADD|SETP)", r[ci["Source"]].strip().lstrip("@!P0123456789 ")):
            a[2] += ins
        for h in stall_cols:
            v = float(r[ci[h]] or 0)
            if v:
                a[3][h] += v
        tot_s += s
        tot_i += ins
    print(f"total samples {tot_s:.0f}, warp instructions {tot_i:.0f}")
    srcs = {}
    for (f, line), (s, ins, fp64, st) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        if 100 * s / tot_s < min_pct and 100 * ins / tot_i < min_pct:
            continue
        if f not in srcs:
            try:
                srcs[f] = open(f"{src_dir}/{f}").read().splitlines()
            except OSError:
                srcs[f] = []
        code = srcs[f][line - 1].strip()[:70] if 0 < line <= len(srcs[f]) else ""
        top = ",".join(f"{k[6:]}:{100*v/max(s,1):.0f}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
        print(f"{f}:{line:4d} smp {100*s/tot_s:5.1f}% inst {100*ins/tot_i:5.1f}% fp64 {100*fp64/max(ins,1):3.0f}% [{top}] {code}")


if __name__ == "__main__":
    main()
