"""Reduces `ncu -i X.ncu-rep --page raw --csv` to one row per profiled launch with the metrics DESIGN.md quotes:
    ncu -i gpurun_out/r2_full_step.ncu-rep --page raw --csv > raw.csv;  python profiles/ncu_summary.py raw.csv > summary.csv"""
import csv
import re
import sys

WANT = [
    ("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
    ("smsp__inst_executed.sum", "warp_inst"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_throughput_pct"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_throughput_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"), ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dyn_smem"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
    ("l1tex__data_bank_conflicts_pipe_lsu.sum", "smem_bank_conflicts"),
]
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
names, units = rows[hdr], rows[hdr + 1]
col = {n: i for i, n in enumerate(names)}
out = csv.writer(sys.stdout)
have = [(m, a) for m, a in WANT if m in col]
out.writerow(["kernel"] + [f"{a} [{units[col[m]]}]" for m, a in have])
for r in rows[hdr + 2:]:
    if len(r) < len(names):
        continue
    k = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "").replace("(anonymous namespace)::", "")
    out.writerow([k] + [r[col[m]] for m, _ in have])
