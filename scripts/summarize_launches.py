#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, mean and share.

    python scripts/summarize_launches.py gpurun_out/launches_r1.csv profiles/r1_launches   # writes .csv (own kernels) + .md
"""
import collections
import csv
import re
import sys


def main(src, dst):
    rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
    hdr = rows[0]
    ki, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("ID")
    own, agg = [], collections.OrderedDict()
    torch_n = torch_ns = 0
    for r in rows[1:]:
        name, ns = r[ki], float(r[vi].replace(",", ""))
        if "at::" in name or "cudnn" in name or "cub::" in name:
            torch_n += 1
            torch_ns += ns
            continue
        short = re.sub(r"\(.*", "", name).replace("void ", "").replace("<unnamed>::", "")
        own.append((r[ii], short, ns))
        a = agg.setdefault(short, [0, 0.0])
        a[0] += 1
        a[1] += ns
    with open(dst + ".csv", "w") as fh:
        fh.write("id,kernel,gpu__time_duration_ns\n")
        for i, k, ns in own:
            fh.write(f'{i},"{k}",{ns:.0f}\n')
    tot = sum(v[1] for v in agg.values())
    with open(dst + ".md", "a") as fh:
        fh.write(f"\n| kernel | launches | total us | mean us | share of own GPU time |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            fh.write(f"| `{k}` | {v[0]} | {v[1] / 1e3:.1f} | {v[1] / v[0] / 1e3:.1f} | {v[1] / tot:.3f} |\n")
        fh.write(f"| **all own kernels** | {len(own)} | {tot / 1e3:.1f} | | 1.000 |\n")
        fh.write(f"\n(torch kernels of the synthetic plate generator, outside the timed region: {torch_n} launches, "
                 f"{torch_ns / 1e3:.0f} us -- not listed)\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
