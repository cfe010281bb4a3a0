"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, mean, share.

usage: python tools/ncu_launch_summary.py gpurun_out/launches.csv > profiles/rNN_launches.md
Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.
"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1], errors="replace")))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        name = r[kn].split("(")[0].replace("<unnamed>::", "").replace("void ", "")[:70]
        v = float(r[mv].replace(",", ""))
        unit = r[mu]
        if unit in ("nsecond", "ns"):
            v /= 1000.0
        elif unit in ("msecond", "ms"):
            v *= 1000.0
        elif unit in ("second", "s"):
            v *= 1e6
        agg.setdefault(name, []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print(f"| kernel | launches | mean us | total us | share |\n|---|---:|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"| `{k}` | {len(v)} | {sum(v) / len(v):.1f} | {sum(v):.0f} | {100 * sum(v) / tot:.1f}% |")
    print(f"\ntotal {tot:.0f} us over {sum(len(v) for v in agg.values())} launches")


if __name__ == "__main__":
    main()
