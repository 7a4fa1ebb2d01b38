"""Headline metrics + per-warp-role stall summary of an ncu report (first launch).  Usage: ncu_summary.py rep [N top instrs]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 15
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines())); hdr, units, r = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum"]
print(r[hdr.index("Kernel Name")][:100])
for k in want:
    if k in hdr: print(f"  {k:80s} {r[hdr.index(k)]} {units[hdr.index(k)]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines())); hdr = rows[1]
isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
ins = [x for x in rows[2:] if len(x) > isamp and x[isamp].isdigit()]
tot = sum(int(x[isamp]) for x in ins); agg = collections.Counter()
for x in ins:
    for i in stall: agg[hdr[i]] += int(x[i] or 0)
print("samples", tot, "instrs", len(ins), agg.most_common(8))
for i, x in sorted(enumerate(ins), key=lambda t: -int(t[1][isamp]))[:top]:
    st = sorted(((hdr[c][6:], int(x[c] or 0)) for c in stall), key=lambda t: -t[1])[:2]
    print(f"{i:5d} {int(x[isamp]):6d} ex={x[iex]:>9s} {x[isrc].strip()[:70]:70s} {st}")
