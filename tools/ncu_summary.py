#!/usr/bin/env python
"""Turns ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches.csv profiles/r1_launches.txt
    python tools/ncu_summary.py full gpurun_out/prof.ncu-rep profiles/r1_ncu_full.txt [profiles/traffic.json]
"""
import collections
import csv
import re
import json
import subprocess
import sys


def kernel_name(full):
    """'void compose_kernel<1>(RenderTables, ...)' -> 'compose_kernel' (template instantiations count as one kernel)"""
    name = full.split("(")[0].strip()
    name = re.sub(r"^void\s+", "", name)
    return re.sub(r"<.*>$", "", name)


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[start]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[start + 1:]:
        name = kernel_name(r[ki])
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v for _, v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised): {src}\n")
        f.write(f"{'kernel':44s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}\n")
        for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k[:44]:44s} {n:8d} {v:12.1f} {v / n:10.1f} {v / tot * 100:6.1f}%\n")
    print(open(dst).read())


METRICS = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "dram__bytes_read.sum",
           "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
           "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "launch__occupancy_limit_registers",
           "launch__occupancy_limit_shared_mem"]


def to_bytes(v, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(v.replace(",", "")) * mult.get(unit, 1)


def full(src, dst, traffic=None):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    stall = [(h.replace("smsp__pcsamp_warps_issue_stalled_", ""), i) for i, h in enumerate(hdr)
             if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h]
    per_kernel = collections.defaultdict(list)
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on: {src}\n")
        for r in data:
            name = kernel_name(r[col["Kernel Name"]])
            f.write(f"\n== {name}\n")
            for m in METRICS:
                if m in col:
                    f.write(f"  {m:64s} {r[col[m]]:>16s} {units[col[m]]}\n")
            tot = sum(float(r[i] or 0) for _, i in stall) or 1.0
            top = sorted(((float(r[i] or 0) / tot * 100, n) for n, i in stall), reverse=True)[:5]
            f.write("  stall reasons (pc sampling): " + ", ".join(f"{n} {p:.1f}%" for p, n in top) + "\n")
            rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
            wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
            per_kernel[name].append(rd + wr)
    print(open(dst).read())
    if traffic:
        js = {k: sum(v) / len(v) for k, v in per_kernel.items()}
        js["_note"] = ("dram__bytes_read.sum + dram__bytes_write.sum per launch (mean over captured launches) from "
                       + src + "; captured on the bench command with --icons 256, scale by icons/256 for other batches")
        json.dump(js, open(traffic, "w"), indent=1)


def traffic(src, dst, icons, skip):
    """src: csv of `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum` over
    tools/one_step.py; the first `skip` launches (the warm-up render) are dropped, the rest is one step."""
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[start]
    ki, mi, vi, ui, ii = (hdr.index(n) for n in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
    per = collections.OrderedDict()
    for r in rows[start + 1:]:
        if int(r[ii]) < skip:
            continue
        name = kernel_name(r[ki])
        d = per.setdefault(name, {"launches": set(), "bytes": 0.0, "us": 0.0})
        d["launches"].add(r[ii])
        v = float(r[vi].replace(",", ""))
        if r[mi].startswith("dram__bytes"):
            d["bytes"] += to_bytes(r[vi], r[ui])
        else:
            d["us"] += v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
    out = {"_note": f"per-step DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum summed over the launches of "
                    f"one resident step) of the bench workload with {icons} icons; ncu times are cold-cache / serialised",
           "icons": int(icons)}
    for k, d in per.items():
        out[k] = {"bytes_per_step": d["bytes"], "launches_per_step": len(d["launches"]), "ncu_us_per_step": d["us"]}
    json.dump(out, open(dst, "w"), indent=1)
    tot = sum(d["us"] for d in per.values())
    for k, d in sorted(per.items(), key=lambda kv: -kv[1]["us"]):
        print(f"{k[:40]:40s} n={len(d['launches']):3d} {d['us']:10.1f} us {d['us'] / tot * 100:5.1f}%  "
              f"{d['bytes'] / 1e6:10.1f} MB")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    elif sys.argv[1] == "traffic":
        traffic(sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5]))
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
