#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    tot = collections.defaultdict(float)
    cnt = collections.Counter()
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        name = re.sub(r"^void |\(anonymous namespace\)::", "", name)
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    print(f"total {T/1e3:.2f} ms over {sum(cnt.values())} launches (ncu: cold cache, serialised)")
    print(f"{'kernel':48s} {'launches':>8s} {'time us':>12s} {'share':>7s}")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f"{k:48s} {cnt[k]:8d} {v:12.1f} {100*v/T:6.2f}%")


if __name__ == "__main__":
    main(sys.argv[1])
