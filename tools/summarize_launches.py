"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel -> markdown share table.

    python tools/summarize_launches.py gpurun_out/launches_r1_b.csv profiles/r1_launches_summary.md
"""
import collections
import csv
import re
import sys


def main(src, dst):
    with open(src) as f:
        lines = [l for l in f if l.startswith('"')]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        val = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        us = val / 1e3 if unit in ("ns", "nsecond") else (val * 1e3 if unit in ("ms", "msecond") else val)
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
        agg[name][0] += 1
        agg[name][1] += us
    tot = sum(v[1] for v in agg.values())
    out = ["| kernel | launches | total us | share | avg us |", "|---|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| {k[:100]} | {v[0]} | {v[1]:.0f} | {100 * v[1] / tot:.1f}% | {v[1] / v[0]:.1f} |")
    out.append(f"| TOTAL | {sum(v[0] for v in agg.values())} | {tot:.0f} | 100% | |")
    text = "\n".join(out) + "\n"
    if dst:
        open(dst, "w").write(text)
    print(text)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
