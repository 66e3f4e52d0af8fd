# 2-GPU visit: slab parity tests (persistent one-launch solve + multi-kernel protocol), slab bench both ways
timeout 400 python -m pytest tests/test_gpu_slab.py -x -q -m gpu 2>&1 | tail -15
for np_ in 1 0; do
  if [ $np_ = 1 ]; then export TB_NO_PERSIST=1; else unset TB_NO_PERSIST; fi
  echo "TB_NO_PERSIST=${TB_NO_PERSIST:-}"
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29539 tools/slab_bench.py --size 2048 2>&1 | tail -1
done
