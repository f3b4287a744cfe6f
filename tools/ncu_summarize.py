"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import collections
import csv
import re
import sys


def main(path, top=40):
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.reader(lines))
    hdr = rows[0]
    ki, vi, ui, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Grid Size")
    agg, tot = collections.OrderedDict(), 0.0
    for r in rows[1:]:
        name = re.sub(r"\(.*", "", r[ki])
        name = re.sub(r"void |<unnamed>::|\(anonymous namespace\)::", "", name)
        v = float(r[vi].replace(",", ""))
        v = v / 1000 if r[ui] == "ns" else (v * 1000 if r[ui] == "ms" else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    print(f"{'us':>10} {'n':>5} {'avg us':>8} {'share':>6}  kernel")
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{v:10.1f} {n:5d} {v / n:8.1f} {100 * v / tot:5.1f}%  {k[:100]}")
    print(f"total {tot:.1f} us over {sum(n for n, _ in agg.values())} launches")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
