#!/usr/bin/env python3
"""Pick the headline metrics out of `ncu -i X.ncu-rep --page raw --csv` output and print a markdown table.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > /tmp/prof.csv
    python scripts/summarize_ncu.py /tmp/prof.csv >> profiles/rN_x.md
"""
import csv
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput % of peak"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared-memory wavefronts % of peak"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
]
STALLS = "smsp__pcsamp_warps_issue_stalled_"


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        print(f"\n### `{name}`\n\n| metric | value | unit |\n|---|---:|---|")
        for key, label in WANT:
            if key in hdr:
                i = hdr.index(key)
                print(f"| {label} (`{key}`) | {r[i]} | {units[i]} |")
        st = []
        for i, h in enumerate(hdr):
            if h.startswith(STALLS) and "not_issued" not in h:
                try:
                    st.append((float(r[i].replace(",", "")), h[len(STALLS):]))
                except ValueError:
                    pass
        tot = sum(v for v, _ in st) or 1.0
        top = sorted(st, reverse=True)[:6]
        print("\nTop warp-stall reasons (pc sampling): " + ", ".join(f"{n} {v / tot:.0%}" for v, n in top))


if __name__ == "__main__":
    main(sys.argv[1])
