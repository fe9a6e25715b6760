"""Summarise an `ncu --metrics gpu__time_duration.sum --csv --log-file X` launch list per kernel.

usage: python profiles/tools/summarise_launches.py gpurun_out/launches.csv > profiles/launches_rNN.md
"""
import collections
import csv
import sys


def main(path):
    lines = open(path).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    rows = list(csv.DictReader(lines[start:]))
    agg = collections.OrderedDict()
    for r in rows:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        ns = float(r["Metric Value"].replace(",", ""))
        if r["Metric Unit"] in ("us", "usecond"):
            ns *= 1e3
        if r["Metric Unit"] in ("ms", "msecond"):
            ns *= 1e6
        k = r["Kernel Name"]
        a = agg.setdefault(k, [0, 0.0, r["Grid Size"], r["Block Size"]])
        a[0] += 1
        a[1] += ns
    tot = sum(a[1] for a in agg.values())
    print(f"launches: {sum(a[0] for a in agg.values())}, summed device time {tot / 1e3:.1f} us "
          f"(cold-cache, serialised by ncu: compare shares, not absolutes)\n")
    print("| kernel | launches | us / launch | share of listed time | grid | block |")
    print("|---|---|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"| `{k[:110]}` | {a[0]} | {a[1] / a[0] / 1e3:.1f} | {100 * a[1] / tot:.1f}% | {a[2]} | {a[3]} |")


if __name__ == "__main__":
    main(sys.argv[1])
