"""Dev tool: per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.

    ncu --metrics gpu__time_duration.sum --clock-control none -c N --csv --log-file L.csv <command>
    python tools/ncu_launch_summary.py L.csv [--md]

(Per-launch times under ncu are cold-cache and serialised: shares, not absolutes, carry over.)
"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    md = "--md" in sys.argv
    hdr = None
    agg = collections.defaultdict(lambda: [0, 0.0])
    with open(path, newline="") as f:
        for r in csv.reader(f):
            if "Kernel Name" in r:
                hdr = r
                continue
            if hdr is None or len(r) != len(hdr):
                continue
            d = dict(zip(hdr, r))
            try:
                v = float(d["Metric Value"].replace(",", ""))
            except ValueError:
                continue
            u = d.get("Metric Unit", "")
            us = v / 1e3 if u.startswith("n") else (v * 1e3 if u.startswith("m") else (v * 1e6 if u.startswith("s") else v))
            k = d["Kernel Name"].split("(")[0][:48]
            agg[k][0] += 1
            agg[k][1] += us
    tot = sum(t for _, t in agg.values()) or 1.0
    if md:
        print("| kernel | launches | total ms | avg us | share |\n|---|---:|---:|---:|---:|")
    for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        if md:
            print(f"| `{k}` | {c} | {t / 1e3:.2f} | {t / c:.2f} | {100 * t / tot:.1f} % |")
        else:
            print(f"{k:48s} {c:7d} launches {t / 1e3:9.2f} ms total {t / c:8.2f} us avg {100 * t / tot:5.1f} %")
    print(f"{'total':48s} {sum(c for c, _ in agg.values()):7d} launches {tot / 1e3:9.2f} ms" if not md else f"| total | {sum(c for c, _ in agg.values())} | {tot / 1e3:.2f} | | |")


if __name__ == "__main__":
    main()
