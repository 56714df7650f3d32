#!/usr/bin/env python3
"""Static SASS mnemonic counts per kernel of the built library (profiles/sass_summary.md, profiles/r2_raw/sass_table.md).

    python scripts/sass_summary.py [path/to/libstitchb200.so] > table.md
"""
import collections
import re
import subprocess
import sys

COLS = ["UTCHMMA", "UTCBAR", "LDTM", "UBLKCP", "UBLKPF", "UTMALDG", "LDGSTS", "SYNCS", "FFMA2", "FMUL2", "FFMA", "LDG.E.*256", "LDL|STL", "MUFU"]


def main(path):
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    cur, counts, total = None, collections.defaultdict(collections.Counter), collections.Counter()
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            cur = cur.replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("void ", "")
            cur = re.sub(r"\(.*", "", cur) or m.group(1)
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if cur and m:
            total[cur] += 1
            for c in COLS:
                if re.match(c, m.group(1)):
                    counts[cur][c] += 1
    print("| kernel | SASS instr | " + " | ".join(c.replace("|", " / ") for c in COLS) + " |")
    print("|---|---:|" + "---:|" * len(COLS))
    for k in sorted(total, key=lambda k: -total[k]):
        print(f"| `{k}` | {total[k]} | " + " | ".join(str(counts[k][c]) if counts[k][c] else "" for c in COLS) + " |")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "image_stitcher_b200/_lib/libstitchb200.so")
