"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / total / share.
Usage: python tools/summarize_launches.py launches.csv "<command line that produced it>" > summary.txt"""
import csv
import sys


def main():
    path = sys.argv[1]
    cmd = sys.argv[2] if len(sys.argv) > 2 else ""
    rows = []
    with open(path, newline="") as fh:
        lines = [l for l in fh if not l.startswith("==")]
    rd = csv.reader(lines)
    hdr = None
    for r in rd:
        if hdr is None:
            if "Kernel Name" in r:
                hdr = {h: i for i, h in enumerate(r)}
            continue
        if len(r) <= max(hdr.values()):
            continue
        if r[hdr["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[hdr["Metric Value"]].replace(",", ""))
        unit = r[hdr["Metric Unit"]]
        us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
        rows.append((r[hdr["Kernel Name"]], us))
    agg = {}
    for name, us in rows:
        a = agg.setdefault(name[:80], [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(v[1] for v in agg.values())
    print(cmd)
    print("launches captured: %d, total %.1f us (cold-cache, serialised: compare shares)" % (len(rows), tot))
    for name, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:50]:
        print("%-82s n=%4d total=%9.1f us avg=%8.2f us %5.1f%%" % (name, c, t, t / c, 100 * t / tot))


if __name__ == "__main__":
    main()
