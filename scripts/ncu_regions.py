"""Summarise an ncu report of assign_tc_kernel: key metrics + stall samples per warp-role region."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
for r in rows[2:3]:
    print(r[hdr.index("Kernel Name")][:70])
    for k in want:
        if k in hdr:
            print(f"  {k:85s} {r[hdr.index(k)]} {units[hdr.index(k)]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]
isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
ins = [r for r in rows[2:] if len(r) > isamp and r[isamp].isdigit()]
marks = [(i, r[isrc].strip()) for i, r in enumerate(ins) if any(k in r[isrc] for k in ("UTCHMMA", "LDTM", "UTMALDG.2D", "UBLKCP", "BAR.SYNC", "UTCBAR"))]
first = lambda key: next(i for i, s in marks if key in s)
last = lambda key: max(i for i, s in marks if key in s)
mma0, ubl, ld0, ld1, tma = first("UTCHMMA"), first("UBLKCP"), first("LDTM"), last("LDTM"), first("UTMALDG.2D")
regions = [("prologue", 0, mma0 - 200), ("mma", mma0 - 200, ubl - 220), ("convert", ubl - 220, ld0 - 180), ("epi_main", ld0 - 180, ld1 + 120),
           ("epi_tail", ld1 + 120, tma - 130), ("producer", tma - 130, tma + 70), ("end", tma + 70, len(ins))]
tot_all = sum(int(r[isamp]) for r in ins)
print("total samples", tot_all, "instructions", len(ins))
for name, a, b in regions:
    a, b = max(a, 0), min(b, len(ins))
    tot, ex, agg = 0, 0, {}
    for r in ins[a:b]:
        tot += int(r[isamp]); ex += int(r[iex])
        for i in stall_cols:
            agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
    top = sorted(agg.items(), key=lambda x: -x[1])[:5]
    print(f"{name:9s} [{a:5d},{b:5d}) samples={tot:6d} ({100*tot/max(tot_all,1):4.1f}%) warp_instr={ex:10d}  {top}")
if len(sys.argv) > 2:
    srt = sorted(ins, key=lambda r: -int(r[isamp]))[:int(sys.argv[2])]
    for r in srt:
        st = sorted(((hdr[i], int(r[i] or 0)) for i in stall_cols), key=lambda x: -x[1])[:2]
        print(f"{int(r[isamp]):6d} ex={r[iex]:>9s} {r[isrc].strip()[:80]:80s} {st}")
