#!/bin/bash
# One GPU-box pass for the round's evidence: GPU parity tests, smoke, the default bench line, the launch list of the
# bench command.  Outputs -> gpurun_out/.
#   gpurun --timeout 900 -- 'bash scripts/gpu_round.sh r2'
tag=${1:-r2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"
timeout 400 python bench.py > gpurun_out/${tag}_bench_gpus1.json 2> gpurun_out/${tag}_bench_gpus1.err; echo "bench rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1; echo "launch list rc=$?"
# (the `ncu --set full` capture of every kernel is a separate, own-kernels-only script: scripts/ncu_all.sh, ~13 GPU-minutes)
tail -3 gpurun_out/${tag}_pytest_gpu.log
