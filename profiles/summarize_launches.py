"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list into a per-kernel table (markdown)."""
import csv
import re
import sys
from collections import defaultdict


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = val / 1e3 if unit in ("ns", "nsecond") else val if unit in ("us", "usecond") else val * 1e3
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        name = re.sub(r"^void ", "", name)
        name = re.sub(r"plb::\(anonymous namespace\)::|plb::<unnamed>::", "", name)
        rows.append((name, us))
    agg = defaultdict(lambda: [0, 0.0])
    for n, us in rows:
        agg[n][0] += 1
        agg[n][1] += us
    total = sum(v[1] for v in agg.values())
    print(f"| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{n[:90]}` | {c} | {t:.0f} | {100 * t / total:.1f}% | {t / c:.1f} |")
    print(f"| **total** | {len(rows)} | {total:.0f} | 100% | |")


if __name__ == "__main__":
    main(sys.argv[1])
