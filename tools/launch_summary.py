"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: totals per kernel and the in-order timeline
of the long launches.  Usage: python tools/launch_summary.py FILE.csv [min_us_for_timeline]"""
import collections
import csv
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr, start = r, i
            break
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    ig, ib = hdr.index("Grid Size"), hdr.index("Block Size")
    out = []
    for r in rows[start + 1:]:
        if len(r) <= iv:
            continue
        v = float(r[iv].replace(",", ""))
        v = v / 1000 if r[iu] == "ns" else v * 1000 if r[iu] == "ms" else v
        out.append((r[ik], v, r[ig], r[ib]))
    return out


def main():
    seq = load(sys.argv[1])
    thr = float(sys.argv[2]) if len(sys.argv) > 2 else 20.0
    tot = sum(v for _, v, _, _ in seq)
    print(f"{len(seq)} launches, {tot:.1f} us summed")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, v, _, _ in seq:
        agg[n[:100]][0] += 1
        agg[n[:100]][1] += v
    ours = sum(t for n, (c, t) in agg.items() if "up3d::" in n)
    print(f"up3d:: kernels: {sum(c for n, (c, t) in agg.items() if 'up3d::' in n)} launches, {ours:.1f} us ({100 * ours / tot:.1f} %)")
    for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:40]:
        print(f"{t:9.1f} us {100 * t / tot:5.1f} % {c:5d} x  {n}")
    print("--- timeline of launches >=", thr, "us")
    cum = 0.0
    for j, (n, v, g, b) in enumerate(seq):
        cum += v
        if v >= thr:
            print(f"{j:5d} {cum:8.1f} {v:7.1f} {g:>16s} {b:>14s} {n[:100]}")


if __name__ == "__main__":
    main()
