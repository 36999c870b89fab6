"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / total / mean, plus per-launch detail.

usage: python tools/launch_summary.py gpurun_out/launches.csv [substring-of-kernel-name-for-detail ...]
"""
import collections
import csv
import re
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hdr]
    ki, vi, gi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size")
    out = []
    for r in rows[hdr + 2:]:
        if len(r) > vi:
            out.append((re.sub(r"\(.*", "", r[ki]), r[gi], float(r[vi].replace(",", "")) / 1e3))
    return out


def main():
    launches = load(sys.argv[1])
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, _, t in launches:
        agg[n[:72]][0] += 1
        agg[n[:72]][1] += t
    tot = sum(v[1] for v in agg.values())
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"{t:10.1f} us {100 * t / tot:5.1f}% {c:5d} x {t / c:8.1f}  {n}")
    print(f"total {tot:.1f} us over {len(launches)} launches")
    for pat in sys.argv[2:]:
        sel = [(g, round(t, 1)) for n, g, t in launches if pat in n]
        print(pat, sel[-12:])


if __name__ == "__main__":
    main()
