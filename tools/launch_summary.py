"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name.
    python tools/launch_summary.py gpurun_out/launches.csv [top_n]
"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    n = 0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1000.0 if unit == "ns" else (v * 1000.0 if unit == "ms" else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        n += 1
    tot = sum(a[1] for a in agg.values())
    ours = sum(a[1] for k, a in agg.items() if "spnb" in k or "k_convs" in k or "k_collide" in k)
    print("launches %d, total %.1f us, libspnb kernels %.1f us (%.1f%%)" % (n, tot, ours, 100 * ours / tot))
    print("%-80s %6s %11s %9s %6s" % ("kernel", "n", "total us", "avg us", "share"))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%-80s %6d %11.1f %9.2f %6.3f" % (k[:80], a[0], a[1], a[1] / a[0], a[1] / tot))


if __name__ == "__main__":
    main()
