"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel.
    python tools/ncu_summary.py gpurun_out/launches.csv > profiles/<name>.txt"""
import collections
import csv
import re
import sys


def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
    print(f"# {path}: {sum(n for n, _ in agg.values())} launches, {tot / 1e3:.3f} ms total (cold-cache, serialised)")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t:12.1f} us {100 * t / tot:5.1f}%  n={n:5d}  avg {t / n:9.1f} us  {k[:100]}")


if __name__ == "__main__":
    main(sys.argv[1])
