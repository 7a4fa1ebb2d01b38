#!/bin/bash
# One GPU-box pass for the round's evidence: GPU parity tests, smoke, the default bench line, the launch list of the
# bench command and one `ncu --set full` capture of every kernel (scripts/prof_all_kernels.py).  Outputs -> gpurun_out/.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh r2'
tag=${1:-r2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"
timeout 400 python bench.py > gpurun_out/${tag}_bench_gpus1.json 2> gpurun_out/${tag}_bench_gpus1.err; echo "bench rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o /tmp/${tag}_all \
    python scripts/prof_all_kernels.py > gpurun_out/${tag}_ncu_all.log 2>&1; echo "ncu full rc=$?"
python scripts/ncu_table.py /tmp/${tag}_all.ncu-rep > gpurun_out/${tag}_all_kernels_ncu_summary.txt 2>&1
ncu -i /tmp/${tag}_all.ncu-rep --page raw --csv > gpurun_out/${tag}_all_kernels_ncu_raw.csv 2>/dev/null
sz=$(stat -c %s /tmp/${tag}_all.ncu-rep 2>/dev/null || echo 0); echo "rep bytes $sz"
if [ "$sz" -lt 45000000 ] && [ "$sz" -gt 0 ]; then cp /tmp/${tag}_all.ncu-rep gpurun_out/; fi
tail -3 gpurun_out/${tag}_pytest_gpu.log
