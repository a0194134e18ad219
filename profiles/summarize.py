#!/usr/bin/env python
"""Turn the raw ncu outputs a gpurun call brought back (gpurun_out/) into the tracked summaries under profiles/.

    python profiles/summarize.py TAG [--launches gpurun_out/launches_TAG.csv] [--rep gpurun_out/prof_TAG.ncu-rep]

  profiles/TAG_launch_list.csv      per-kernel launches / total / average / share of the `--metrics gpu__time_duration.sum` pass
  profiles/TAG_ncu_full_summary.txt the metrics of the `--set full` capture that the roofline discussion in DESIGN.md cites
  profiles/ncu_traffic.json         DRAM bytes per launch of the E and M kernels (bench.py's roofline.traffic)
"""
import argparse, collections, csv, io, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
L2_XBAR_PEAK_TBS = 21.07   # profiles/probes/l2_cap_probe.out
KEEP = re.compile(r"^(gpu__time_duration\.sum|dram__bytes_(read|write)\.sum|lts__t_sector_hit_rate\.pct|l1tex__t_sector_hit_rate\.pct|"
                  r"lts__throughput\.avg\.pct_of_peak_sustained_elapsed|l1tex__throughput\.avg\.pct_of_peak_sustained_elapsed|"
                  r"sm__throughput\.avg\.pct_of_peak_sustained_elapsed|sm__warps_active\.avg\.pct_of_peak_sustained_active|"
                  r"launch__registers_per_thread|smsp__inst_executed\.sum|launch__grid_size|smsp__issue_active\.avg\.pct_of_peak_sustained_active|"
                  r"derived__lts__lts2xbar_bytes\.sum\.per_second|l1tex__data_pipe_lsu_wavefronts\.avg\.pct_of_peak_sustained_elapsed|"
                  r"smsp__average_warps_issue_stalled_(long_scoreboard|wait|no_instruction|branch_resolving|barrier)_per_issue_active\.ratio|"
                  r"sm__pipe_tensor_cycles_active\.avg\.pct_of_peak_sustained_active|launch__shared_mem_per_block_dynamic)$")


def launch_list(path, out, note):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    h = rows[0]
    agg = collections.OrderedDict()
    for r in rows[1:]:
        d = dict(zip(h, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"^void ", "", d["Kernel Name"])
        name = re.sub(r"\(.*$", "", name)
        name = re.sub(r"<.*$", "", name)
        v = float(d["Metric Value"].replace(",", ""))
        if d.get("Metric Unit", "ns") in ("us", "usecond"):
            v *= 1e3
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as fh:
        fh.write(f"# ncu launch list, {note}\n# cold-cache, serialised launches: compare SHARES, not absolutes\n")
        fh.write(f"# total {tot / 1e6:.3f} ms over {sum(a[0] for a in agg.values())} launches\nkernel,launches,total_ms,avg_us,share_pct\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            fh.write(f"{k},{a[0]},{a[1] / 1e6:.3f},{a[1] / a[0] / 1e3:.1f},{100 * a[1] / tot:.1f}\n")
    return out


def full_summary(rep, out, note):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, units = rows[0], rows[1]
    traffic = {}
    with open(out, "w") as fh:
        fh.write(f"# ncu --set full --clock-control none, {note}\n")
        for r in rows[2:]:
            d = dict(zip(h, r))
            fh.write(f"---- {d['Kernel Name']}\n")
            for name, u in zip(h, units):
                if KEEP.match(name):
                    fh.write(f"   {name:88s} {d[name]:>18s} {u}\n")
            kern = "estep_kernel" if "estep" in d["Kernel Name"] else ("mstep_kernel" if "mstep_kernel" in d["Kernel Name"] else None)
            if kern and kern not in traffic:
                def to_bytes(key):
                    v, u = float(d[key].replace(",", "")), units[h.index(key)]
                    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
                def num(key, scale=None):
                    v, u = float(d[key].replace(",", "")), units[h.index(key)]
                    return v * (scale or {}).get(u, 1.0)
                dram = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
                us = num("gpu__time_duration.sum", {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3})
                xbar = num("derived__lts__lts2xbar_bytes.sum.per_second", {"Tbyte": 1.0, "Gbyte": 1e-3, "Tbyte/second": 1.0, "Gbyte/second": 1e-3})
                traffic[kern] = {"dram_bytes_per_launch": dram, "capture_launch_us": us, "dram_gbs_real": dram / (us * 1e-6) / 1e9,
                                 "l2_xbar_tbs": xbar, "l2_xbar_frac": xbar / L2_XBAR_PEAK_TBS, "l2_xbar_peak_tbs": L2_XBAR_PEAK_TBS,
                                 "l2_xbar_peak_source": "profiles/probes/l2_cap_probe.out: best L2 -> SM row-gather rate measured on this B200 (16 MB table, 64 warps/SM)",
                                 "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                                 "source": f"{os.path.relpath(out, ROOT)} (ncu --set full, one full-occupancy launch, 50 resident restarts)"}
    return traffic


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("tag")
    ap.add_argument("--launches")
    ap.add_argument("--rep")
    ap.add_argument("--launch-note", default="python bench.py --steps 1 --warmup 0 --no-cpu-baseline; --metrics gpu__time_duration.sum --clock-control none")
    ap.add_argument("--rep-note", default="the E and M kernels of one full-occupancy iteration (50 resident restarts, k = 8, config 2: 10,000 cells x 39,930 loci, nnz 6,441,922)")
    a = ap.parse_args()
    if a.launches:
        print(launch_list(a.launches, os.path.join(ROOT, "profiles", f"{a.tag}_launch_list.csv"), f"snapshot {a.tag} ({a.launch_note})"))
    if a.rep:
        t = full_summary(a.rep, os.path.join(ROOT, "profiles", f"{a.tag}_ncu_full_summary.txt"), f"snapshot {a.tag}: {a.rep_note}")
        if t:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w") as fh:
                json.dump(t, fh, indent=1)
        print(t)
