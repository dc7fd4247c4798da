"""Summarise an `ncu --metrics gpu__time_duration.sum,<pct metrics...> --csv` run: per kernel the launches, total ms, share and the
time-weighted mean of every other metric (issue utilisation, tensor-pipe activity, DRAM throughput ...)."""
import collections
import csv
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    idi, ki, mi, vi, ui = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    launches = collections.OrderedDict()                  # launch id -> (kernel, {metric: value})
    for r in rows[hi + 1:]:
        if len(r) <= vi or not r[vi]:
            continue
        name = r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        d = launches.setdefault(r[idi], (name, {}))[1]
        v = float(r[vi].replace(",", ""))
        if r[mi] == "gpu__time_duration.sum":
            v = v / 1e6 if r[ui] == "ns" else (v / 1e3 if r[ui] == "us" else v)
        d[r[mi]] = v
    agg = collections.OrderedDict()
    metrics = []
    for name, d in launches.values():
        ms = d.get("gpu__time_duration.sum", 0.0)
        a = agg.setdefault(name, {"n": 0, "ms": 0.0})
        a["n"] += 1
        a["ms"] += ms
        for m, v in d.items():
            if m == "gpu__time_duration.sum":
                continue
            if m not in metrics:
                metrics.append(m)
            a[m] = a.get(m, 0.0) + v * ms
    tot = sum(a["ms"] for a in agg.values())
    short = [m.split(".")[0].replace("smsp__", "").replace("sm__", "").replace("gpu__", "")[:22] for m in metrics]
    print("# total kernel time %.1f ms over %d launches; metric columns are time-weighted means" % (tot, len(launches)))
    print("%-56s %6s %9s %6s " % ("kernel", "n", "total_ms", "share") + " ".join("%22s" % s for s in short))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])[:top]:
        print("%-56s %6d %9.2f %5.1f%% " % (k[:56], a["n"], a["ms"], 100 * a["ms"] / tot) +
              " ".join("%22.1f" % (a.get(m, 0.0) / a["ms"] if a["ms"] else 0.0) for m in metrics))


if __name__ == "__main__":
    main(sys.argv[1])
