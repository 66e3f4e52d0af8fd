#!/usr/bin/env bash
# Multi-GPU visit (gpurun --gpus N): slab parity tests at every rank count the box offers, the slab solve in its
# variants (tools/slab_bench.py), a timeline of the one-launch solve, and bench.py under torch.distributed.run.
# Usage: bash tools/multi_gpu.sh <tag> [quick]
set -u
TAG=${1:-r02}
QUICK=${2:-}
OUT=gpurun_out
mkdir -p $OUT
NG=$(nvidia-smi -L | wc -l)
RUN="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
echo "== $NG GPUs; slab parity tests (T7)"
timeout 900 python -m pytest tests/test_gpu_slab.py -x -q -m gpu 2>&1 | tail -3 | tee $OUT/pytest_slab_${TAG}_${NG}gpu.log
port=29540
for P in 2 4 8; do
  [ $P -le $NG ] || continue
  for size in 2048 4096 1024; do
    for variant in "default" "TB_SLAB_SYNC=0" "TB_SLAB_SYNC=1" "TB_SLAB_SYSFENCE=1" "TB_NO_PERSIST=1"; do
      [ -n "$QUICK" ] && [ "$variant" = "TB_NO_PERSIST=1" ] && continue
      port=$((port + 1))
      envs=""; [ "$variant" != "default" ] && envs="$variant"
      line=$(env $envs timeout 120 $RUN --nproc-per-node $P --master-port $port tools/slab_bench.py --size $size 2>&1 | tail -1)
      echo "P=$P size=$size $variant: $line" | tee -a $OUT/slab_${TAG}_${NG}gpu.txt | cut -c1-220
    done
  done
  for size in 2048 1024; do
    for mode in 1 2; do
      port=$((port + 1))
      TB_SLAB_SYNC=$mode TB_SLAB_TIMELINE=$OUT/timeline_${TAG}_P${P}_${size}_mode${mode} timeout 120 $RUN --nproc-per-node $P --master-port $port tools/slab_bench.py --size $size --iters 60 2>&1 | tail -1 | cut -c1-200
    done
  done
done
[ -n "$QUICK" ] && exit 0
echo "== bench.py under torch.distributed.run at every rank count"
for P in 2 4 8; do
  [ $P -le $NG ] || continue
  port=$((port + 1))
  timeout 600 $RUN --nproc-per-node $P --master-port $port bench.py --gpus $P --steps 5 --warmup 3 --no-hmc 2>&1 | tail -1 | tee $OUT/bench_${TAG}_${P}gpu.json | cut -c1-300
done
ls -la $OUT | tail -12
