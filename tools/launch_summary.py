"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, average, share."""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1000.0 if unit in ("nsecond", "ns") else v * 1000.0 if unit in ("msecond", "ms") else v
        name = row["Kernel Name"]
        m = re.match(r"(?:void )?(?:b200::)?(\w+)(<[^(]*>)?", name)
        key = (m.group(1) + (m.group(2) or ""))[:70] if m else name[:70]
        agg[key][0] += 1
        agg[key][1] += v
        tot += v
    print(f"total {tot:.1f} us over {sum(n for n, _ in agg.values())} launches")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:72s} n={n:4d} total={t:9.1f}us avg={t / n:8.2f}us share={100 * t / tot:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
