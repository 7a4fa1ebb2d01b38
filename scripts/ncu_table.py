"""One line per kernel launch of an ncu report: duration, DRAM bytes and throughput, pipe utilisation, issue slots,
registers, achieved occupancy.  Usage: python scripts/ncu_table.py gpurun_out/r2_all.ncu-rep > profiles/r2_all_kernels_ncu_summary.txt"""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
cols = [("gpu__time_duration.sum", "dur"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor%"),
        ("sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "alu%"),
        ("sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active", "fma%"),
        ("sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active", "lsu%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("lts__t_sector_hit_rate.pct", "l2hit%"), ("launch__registers_per_thread", "regs"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("launch__grid_size", "grid"), ("launch__block_size", "block")]
idx = {k: hdr.index(k) for k, _ in cols if k in hdr}
ik = hdr.index("Kernel Name")
print(f"# ncu --set full --clock-control none, one launch per row ({rep}); units: " +
      ", ".join(f"{n}[{units[idx[k]]}]" for k, n in cols if k in idx))
print(f"{'kernel':60s} " + " ".join(f"{n:>10s}" for k, n in cols if k in idx))
for r in rows[2:]:
    if len(r) <= ik:
        continue
    name = r[ik].split("(")[0][-60:]
    print(f"{name:60s} " + " ".join(f"{r[idx[k]]:>10s}" for k, _ in cols if k in idx))
