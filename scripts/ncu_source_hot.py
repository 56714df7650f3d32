"""Hottest instructions and stall mix per kernel of an `ncu --page source --csv --print-source sass` export.

    python scripts/ncu_source_hot.py gpurun_out/c17_src_all.csv [kernel-index] [top-n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[2]) if len(sys.argv) > 2 else -1
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 30
kern, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        kern.append(cur)
    elif r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and r:
        cur["rows"].append(r)
seen = set()
for ki, k in enumerate(kern):
    if k["name"] in seen or (which >= 0 and ki != which):
        continue
    seen.add(k["name"])
    h = k["hdr"]
    iS, iE, isrc = h.index("# Samples"), h.index("Instructions Executed"), h.index("Source")
    cols = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
    tot = sum(int(r[iS]) for r in k["rows"])
    print("=====", ki, k["name"], "samples", tot, "sass instructions", len(k["rows"]),
          "warp-instructions", sum(int(r[iE]) for r in k["rows"]))
    st = {h[i][6:]: sum(int(r[i]) for r in k["rows"]) for i in cols}
    print("   ", {n: round(100 * v / tot, 1) for n, v in sorted(st.items(), key=lambda x: -x[1])[:8]})
    top = sorted(range(len(k["rows"])), key=lambda i: -int(k["rows"][i][iS]))[:topn]
    for i in sorted(top):
        r = k["rows"][i]
        ss = sorted([(int(r[c]), h[c][6:]) for c in cols], reverse=True)[:2]
        print(f"  [{i:5d}] {100 * int(r[iS]) / tot:5.1f}%  exec {r[iE]:>9s}  {r[isrc].strip()[:64]:64s} {ss}")
