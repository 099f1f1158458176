"""Summarises `ncu --page raw --csv` exports (bench/ncu_families.sh) as a markdown table: one row per captured launch.
Usage: python bench/ncu_summary.py profiles/r2_ncu/*.csv > profiles/r2_ncu_families.md"""
import csv
import os
import sys

COLS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU %"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("smsp__inst_executed.sum", "warp instr")]


def main(paths):
    print("| capture | kernel | " + " | ".join(c[1] for c in COLS) + " |")
    print("|---|---|" + "---|" * len(COLS))
    for f in paths:
        rows = list(csv.reader(open(f)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            cells = []
            for key, _ in COLS:
                if key not in idx:
                    cells.append("-")
                    continue
                v, u = r[idx[key]], units[idx[key]]
                try:
                    x = float(v.replace(",", ""))
                    v = f"{x:.3f}" if x < 1000 else f"{x:.0f}"
                except ValueError:
                    pass
                cells.append(f"{v} {u}".strip() if ("time" in key or "bytes" in key) else v)
            print(f"| {os.path.basename(f).replace('ncu_', '').replace('_raw.csv', '')} | `{r[idx['Kernel Name']][:48]}` | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main(sys.argv[1:])
