"""Executed warp instructions per CUDA source line of one kernel, from an ncu report captured with --import-source on.

    ncu -i gpurun_out/prof_r1i.ncu-rep --page source --print-source cuda,sass --csv > /tmp/src_mixed.csv
    python profiles/sass_lines.py /tmp/src_mixed.csv 'estep_kernel<(int)8>' > profiles/r1i_estep_lines.txt

Shares are what matters: the report holds every captured launch of the kernel and inlined lines repeat per call site.
"""
import collections
import csv
import re
import sys

CTL = ("BRA", "BSSY", "BSYNC", "ISETP", "PLOP3", "BREAK")
ESS = ("FADD2", "FADD", "FMUL", "LDG")


def main(path, kernel, top=30):
    fn = None
    cur = None
    text = {}
    per = collections.defaultdict(collections.Counter)
    for r in csv.reader(open(path)):
        if not r:
            continue
        if r[0] == "Function Name":
            fn = r[1]
            continue
        if r[0] in ("File Path", "Line No") or fn is None or kernel not in fn:
            continue
        if r[0] != "":
            cur = int(r[0])
            text[cur] = r[1].strip()
            continue
        m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", r[3]) if len(r) > 7 else None
        if not m:
            continue
        try:
            n = int(float(r[7] or 0))
        except ValueError:
            continue
        c = per[cur]
        c["n"] += n
        c["ctl"] += n if m.group(1) in CTL else 0
        c["ess"] += n if m.group(1) in ESS else 0
    tot = sum(c["n"] for c in per.values())
    print(f"{kernel}: {tot / 1e6:.1f} M warp instructions over the captured launches; control (BRA/BSSY/BSYNC/ISETP/PLOP3) "
          f"{100 * sum(c['ctl'] for c in per.values()) / tot:.1f} %, loads + f32 adds / multiplies {100 * sum(c['ess'] for c in per.values()) / tot:.1f} %")
    for ln, c in sorted(per.items(), key=lambda kv: -kv[1]["n"])[:top]:
        print(f"  line {ln:4d} {100 * c['n'] / tot:5.1f} %  control {100 * c['ctl'] / tot:4.1f} %  essential {100 * c['ess'] / tot:4.1f} %  | {text.get(ln, '')[:120]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
