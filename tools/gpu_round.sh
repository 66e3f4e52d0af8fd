#!/usr/bin/env bash
# One GPU visit: parity tests, smoke, bench, ncu launch list + full capture of the top kernels.
# Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh [tag]
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $OUT/smi_$TAG.txt 2>&1
echo "== pytest -m gpu"; python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee $OUT/pytest_gpu_$TAG.log
echo "== smoke"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke_$TAG.log
echo "== bench"; python bench.py --steps 10 --warmup 3 2>&1 | tail -3 | tee $OUT/bench_$TAG.json
echo "== bench reference arm"; python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee $OUT/bench_ref_$TAG.json
echo "== probe"; python tools/probe.py 2>&1 | tail -12 | tee $OUT/probe_$TAG.log
echo "== ncu launch list"
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 400 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_launch_$TAG.log 2>&1
tail -2 $OUT/ncu_launch_$TAG.log
echo "== ncu full"
ncu --set full --clock-control none --import-source on -k regex:'dslash_kernel|axpy_norm|xpay|resident' -s 40 -c 8 \
    -f -o $OUT/prof_$TAG python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_full_$TAG.log 2>&1
tail -2 $OUT/ncu_full_$TAG.log
ls -la $OUT
