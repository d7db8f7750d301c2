#!/usr/bin/env python
"""Table of an ncu --csv metrics log (tools/ncu_kernels_run.py): per kernel the LONGEST launch (the finest level) with its
duration, DRAM bytes read / written, DRAM throughput in GB/s and as a fraction of the measured copy peak, registers,
achieved occupancy and issue-slot utilisation.
  METRICS = gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread,
            sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active
  tools/summarize_ncu.py log.csv [peak_GBs]"""
import collections
import csv
import json
import os
import re
import sys

METRICS = ("gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread,"
           "sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active")


def to_bytes(v, unit):
    f = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(v) * f.get(unit, 1.0)


def to_us(v, unit):
    f = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}
    return float(v) * f.get(unit, 1.0)


def main():
    path = sys.argv[1]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    peak = float(sys.argv[2]) if len(sys.argv) > 2 else json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))["hbm_gbs"]
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]
    iid, ik, im, iu, iv = h.index("ID"), h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Unit"), h.index("Metric Value")
    launches = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= iv or not r[iid].strip().isdigit():
            continue
        d = launches.setdefault(int(r[iid]), {"name": re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("<unnamed>::", "")})
        v = r[iv].replace(",", "")
        try:
            if r[im].startswith("gpu__time_duration"):
                d["us"] = to_us(v, r[iu])
            elif r[im].startswith("dram__bytes_read"):
                d["rd"] = to_bytes(v, r[iu])
            elif r[im].startswith("dram__bytes_write"):
                d["wr"] = to_bytes(v, r[iu])
            elif r[im].startswith("launch__registers"):
                d["regs"] = float(v)
            elif r[im].startswith("sm__warps_active"):
                d["occ"] = float(v)
            elif r[im].startswith("smsp__issue_active"):
                d["issue"] = float(v)
        except ValueError:
            pass
    best, count = {}, collections.Counter()
    for d in launches.values():
        if "us" not in d:
            continue
        count[d["name"]] += 1
        if d["name"] not in best or d["us"] > best[d["name"]]["us"]:
            best[d["name"]] = d
    print(f"# per kernel: the longest of its launches (ncu, cold caches, serialised); DRAM peak = {peak:.1f} GB/s (measured copy)")
    print(f"{'kernel':44s} {'launches':>8s} {'us':>9s} {'rd MB':>9s} {'wr MB':>9s} {'GB/s':>8s} {'of peak':>8s} {'regs':>5s} {'occ %':>6s} {'issue %':>8s}")
    for name, d in sorted(best.items(), key=lambda kv: -kv[1]["us"]):
        tot = d.get("rd", 0.0) + d.get("wr", 0.0)
        gbs = tot / (d["us"] * 1e-6) / 1e9 if d["us"] > 0 else 0.0
        print(f"{name[:44]:44s} {count[name]:8d} {d['us']:9.1f} {d.get('rd', 0) / 1e6:9.1f} {d.get('wr', 0) / 1e6:9.1f} {gbs:8.0f} "
              f"{gbs / peak:8.2f} {d.get('regs', 0):5.0f} {d.get('occ', 0):6.1f} {d.get('issue', 0):8.1f}")


if __name__ == "__main__":
    main()
