"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total, average, share.
usage: python tools/summarize_launches.py file.csv [--seq]   (--seq also prints the launch sequence)"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        name = name.replace("xkv::", "").replace("void ", "")
        val = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = val / 1000.0 if unit in ("ns", "nsecond") else (val if unit in ("us", "usecond") else val * 1000.0)
        rows.append((name, us, r.get("Grid Size", ""), r.get("Block Size", "")))
    total = sum(u for _, u, _, _ in rows)
    agg = OrderedDict()
    for n, u, _, _ in rows:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += u
    print(f"total {total / 1000:.2f} ms over {len(rows)} launches\n")
    print("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|")
    for n, (c, u) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{n}` | {c} | {u:.0f} | {u / c:.1f} | {100 * u / total:.1f}% |")
    if "--seq" in sys.argv:
        for i, (n, u, g, b) in enumerate(rows):
            print(f"{i:4d} {u:9.1f} us  {n}  grid {g}")


if __name__ == "__main__":
    main()
