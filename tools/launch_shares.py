#!/usr/bin/env python3
"""Per-kernel totals and shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: python tools/launch_shares.py launches.csv"""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    i = [k for k, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[i]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot, cnt = collections.OrderedDict(), collections.Counter()
    for r in rows[i + 1:]:
        if len(r) <= mv:
            continue
        name = r[kn].split("(")[0].replace("void ", "").replace("gsb::", "")
        tot[name] = tot.get(name, 0.0) + float(r[mv])
        cnt[name] += 1
    T = sum(tot.values())
    print(f"{path}: {T / 1e6:.3f} ms of kernel time over {sum(cnt.values())} launches (cold-cache, serialised: compare shares)")
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        print(f"  {k:45s} n={cnt[k]:4d}  total {v / 1e6:8.3f} ms  avg {v / cnt[k] / 1e3:8.1f} us  share {100 * v / T:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
