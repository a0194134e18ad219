"""Dynamic opcode mix of the kernels in an ncu report: executed warp instructions and stall samples per SASS opcode.

    ncu -i gpurun_out/prof_r1i.ncu-rep --page source --csv > /tmp/src.csv
    python profiles/sass_mix.py /tmp/src.csv > profiles/r1i_sass_mix.txt
"""
import collections
import csv
import re
import sys

GROUPS = {
    "essential (row gathers, adds, multiplies)": ("FADD2", "FADD", "FMUL", "LDG.128", "LDG.64", "FFMA", "DFMA"),
    "control flow": ("BRA", "BSSY", "BSYNC", "ISETP", "PLOP3", "BREAK", "EXIT", "UISETP"),
    "integer / address / decode": ("IMAD", "LOP3", "IADD3", "SHF", "LEA", "VIADD", "MOV", "UMOV", "I2FP", "SEL", "LDC", "LDCU", "LDG.32"),
}


def main(path):
    rows = list(csv.reader(open(path)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    seen = set()
    for k, start in enumerate(starts):
        name = rows[start][1]
        if name in seen:
            continue
        seen.add(name)
        end = starts[k + 1] if k + 1 < len(starts) else len(rows)
        hdr = rows[start + 1]
        ci = {h: j for j, h in enumerate(hdr)}
        ops, stall = collections.Counter(), collections.Counter()
        for r in rows[start + 2:end]:
            if len(r) < len(hdr):
                continue
            m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", r[ci["Source"]])
            if not m:
                continue
            op, mod = m.group(1), m.group(2)
            if op in ("LDG", "LDS", "STG", "STS", "LDGSTS"):
                w = re.search(r"\.(128|64)", mod)
                op += "." + (w.group(1) if w else "32")
            try:
                ops[op] += int(float(r[ci["Instructions Executed"]] or 0))
                stall[op] += int(float(r[ci["# Samples"]] or 0))
            except ValueError:
                pass
        total, samples = sum(ops.values()), max(sum(stall.values()), 1)
        print(f"{name}: {total / 1e6:.1f} M warp instructions executed, {samples} stall samples")
        for g, members in GROUPS.items():
            print(f"  {g}: {100 * sum(ops[o] for o in members) / total:.1f} % of the instructions, {100 * sum(stall[o] for o in members) / samples:.1f} % of the samples")
        for op, n in ops.most_common(24):
            print(f"     {op:10s} {n / 1e6:8.2f} M  {100 * n / total:5.1f} %   samples {100 * stall[op] / samples:5.1f} %")
        print()


if __name__ == "__main__":
    main(sys.argv[1])
