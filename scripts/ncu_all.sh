#!/bin/bash
# `ncu --set full` of every kernel of libequss_b200.so once (scripts/prof_all_kernels.py), own kernels only.
#   gpurun --timeout 1500 -- 'bash scripts/ncu_all.sh r2'
tag=${1:-r2}
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled \
    -k regex:equss:: -f -o /tmp/${tag}_all python scripts/prof_all_kernels.py > gpurun_out/${tag}_ncu_all.log 2>&1; echo "ncu full rc=$?"
python scripts/ncu_table.py /tmp/${tag}_all.ncu-rep > gpurun_out/${tag}_all_kernels_ncu_summary.txt 2>&1
ncu -i /tmp/${tag}_all.ncu-rep --page raw --csv > gpurun_out/${tag}_all_kernels_ncu_raw.csv 2>/dev/null
sz=$(stat -c %s /tmp/${tag}_all.ncu-rep 2>/dev/null || echo 0); echo "rep bytes $sz"
if [ "$sz" -lt 45000000 ] && [ "$sz" -gt 0 ]; then cp /tmp/${tag}_all.ncu-rep gpurun_out/; fi
wc -l gpurun_out/${tag}_all_kernels_ncu_summary.txt
