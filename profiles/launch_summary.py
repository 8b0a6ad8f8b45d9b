#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: launches, total and
mean device time, share.  Usage: python profiles/launch_summary.py launches.csv [out.md] [title]"""
import collections
import csv
import re
import sys


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[1:]:
        v = float(r[iv].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1.0)
        name = re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("<unnamed>::", "")
        name = re.sub(r"at::native::|at::", "", name)[:90]
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    title = sys.argv[3] if len(sys.argv) > 3 else sys.argv[1]
    out = [f"# ncu launch list: {title}\n", f"{sum(cnt.values())} launches, {T / 1e3:.3f} ms of device time "
           "(cold-cache, serialised by ncu: compare shares, not absolutes)\n",
           "| kernel | launches | total us | mean us | share |", "|---|---|---|---|---|"]
    for k, v in tot.most_common():
        out.append(f"| `{k}` | {cnt[k]} | {v:.1f} | {v / cnt[k]:.1f} | {100 * v / T:.1f}% |")
    text = "\n".join(out) + "\n"
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
