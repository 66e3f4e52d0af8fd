#!/usr/bin/env bash
# The 8-GPU visit (gpurun --gpus 8; charged 8x): slab parity at 4 and 8 ranks, the slab solve of 2048^2 and 4096^2 in
# its forms at 8 and 4 ranks with a timeline, and bench.py at 8 ranks.  Usage: bash tools/multi_gpu8.sh <tag>
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
NG=$(nvidia-smi -L | wc -l)
RUN="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi topo -m > $OUT/topo_${TAG}.txt 2>&1
echo "== $NG GPUs; slab parity tests (T7) at 4 and 8 ranks"
timeout 900 python -m pytest tests/test_gpu_slab.py -x -q -m gpu -k "${SLAB_TESTS:-P4 or P8}" 2>&1 | tail -3 | tee $OUT/pytest_slab_${TAG}_${NG}gpu.log
port=29600
run() {  # P size extra-env...
  local P=$1 size=$2; shift 2
  port=$((port + 1))
  local line
  line=$(env "$@" timeout 150 $RUN --nproc-per-node $P --master-port $port tools/slab_bench.py --size $size 2>&1 | tail -1)
  echo "P=$P size=$size $*: $line" | tee -a $OUT/slab_${TAG}_${NG}gpu.txt | cut -c1-230
}
run 8 2048 TB_X=0
run 8 2048 TB_SLAB_SYNC=0
run 8 2048 TB_SLAB_SYSFENCE=1
run 8 4096 TB_X=0
run 8 1024 TB_X=0
run 4 2048 TB_X=0
run 4 4096 TB_X=0
port=$((port + 1))
TB_SLAB_TIMELINE=$OUT/timeline_${TAG}_P8_2048 timeout 150 $RUN --nproc-per-node 8 --master-port $port tools/slab_bench.py --size 2048 --iters 60 2>&1 | tail -1 | cut -c1-200
echo "== bench.py at 8 ranks"
port=$((port + 1))
timeout 900 $RUN --nproc-per-node 8 --master-port $port bench.py --gpus 8 --steps 10 --warmup 3 2>&1 | tail -1 | tee $OUT/bench_${TAG}_8gpu.json | cut -c1-300
ls -la $OUT | tail -8
