#!/usr/bin/env python
"""Per-kernel summary (launches, average, total, share) of an ncu launch list made with
`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file <csv> <command>`.

    python tools/summarize_launches.py profiles/r02_final_launches.csv > profiles/r02_final_launches_summary.md
"""
import collections
import csv
import sys


def main(path):
    rows = list(csv.DictReader(l for l in open(path) if not l.startswith("==")))
    per = collections.defaultdict(list)
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v_us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        name = r["Kernel Name"]
        per[name].append(v_us)
    total = sum(sum(v) for v in per.values())
    print("| kernel | launches | avg us | total ms | share |")
    print("|---|---|---|---|---|")
    for name, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
        short = name if len(name) <= 60 else name[:57] + "..."
        print("| {} | {} | {:.1f} | {:.2f} | {:.2f}% |".format(short, len(v), sum(v) / len(v), sum(v) / 1e3, 100 * sum(v) / total))


if __name__ == "__main__":
    main(sys.argv[1])
