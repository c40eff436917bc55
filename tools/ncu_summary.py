"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel:
    python tools/ncu_summary.py gpurun_out/launches.csv > profiles/<name>.txt
or pull the DRAM traffic of one kernel out of an `ncu --set full ... --page raw --csv` export, as the JSON
bench.py reads for roofline.traffic:
    python tools/ncu_summary.py --traffic raw.csv "<kernel substring>" "<label>" <algorithmic_bytes> > profiles/r02_traffic.json
"""
import collections
import csv
import json
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


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def traffic(path, kernel, label, algo_bytes):
    """raw page: one row per launch, one column per metric (two header rows: names, units)"""
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    names, units = rows[0], rows[1]
    col = {n: i for i, n in enumerate(names)}
    rd, wr, kn = col["dram__bytes_read.sum"], col["dram__bytes_write.sum"], col["Kernel Name"]
    hits = [r for r in rows[2:] if kernel in r[kn]]
    if not hits:
        raise SystemExit(f"no launch of {kernel!r} in {path}")
    per = [to_bytes(r[rd], units[rd]) + to_bytes(r[wr], units[wr]) for r in hits]
    dur = col.get("gpu__time_duration.sum")
    out = {"kernel": label, "bytes_per_launch": sum(per) / len(per), "launches_captured": len(per),
           "algorithmic_bytes_per_launch": float(algo_bytes), "source": f"profiles/{path.split('/')[-1]} (ncu --set full, this round)"}
    if dur is not None:
        out["duration_us_under_ncu"] = sum(float(r[dur].replace(",", "")) for r in hits) / len(hits) * \
            {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(units[dur], 1.0)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "--traffic":
        traffic(sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5])
    else:
        main(sys.argv[1])
