"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total ms, share."""
import collections
import csv
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    n = 0
    for r in rows[hi + 1:]:
        if len(r) <= vi or not r[vi]:
            continue
        v = float(r[vi].replace(",", ""))
        ms = v / 1e6 if r[ui] == "ns" else (v / 1e3 if r[ui] == "us" else v)
        a = agg.setdefault(r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", ""), [0, 0.0])
        a[0] += 1
        a[1] += ms
        n += 1
    tot = sum(a[1] for a in agg.values())
    print("# total kernel time %.1f ms over %d launches" % (tot, n))
    print("%-70s %9s %10s %7s" % ("kernel", "launches", "total_ms", "share"))
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%-70s %9d %10.2f %6.1f%%" % (k[:70], c, t, 100 * t / tot))


if __name__ == "__main__":
    main(sys.argv[1])
