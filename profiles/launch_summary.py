#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total
device time and share (cold-cache, serialised: compare shares, not absolutes)."""
import collections
import csv
import sys


def main(path, out, title):
    rows = list(csv.reader(open(path)))
    hdr, data = None, []
    for r in rows:
        if r and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            data.append(dict(zip(hdr, r)))
    tot = collections.Counter()
    cnt = collections.Counter()
    for d in data:
        name = d["Kernel Name"].split("(")[0].replace("void ", "")
        tot[name] += float(d["Metric Value"]) / 1e3
        cnt[name] += 1
    total = sum(tot.values())
    lines = [f"# {title}", f"# total {total / 1e3:.2f} ms over {len(data)} launches", "kernel,launches,total_us,share"]
    for k, v in tot.most_common():
        lines.append(f"{k},{cnt[k]},{v:.1f},{v / total:.4f}")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:14]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else sys.argv[1])
