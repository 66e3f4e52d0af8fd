# 2-GPU visit: slab_bench at three sizes, multi-kernel form (TB_NO_PERSIST=1) and one-launch solve, plus rows-per-thread variants
for size in 1024 512; do
for np_ in 1 0; do
  if [ $np_ = 1 ]; then export TB_NO_PERSIST=1; else unset TB_NO_PERSIST; fi
  echo "size $size TB_NO_PERSIST=${TB_NO_PERSIST:-}"
  timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29539 tools/slab_bench.py --size $size 2>&1 | tail -1
done
done
unset TB_NO_PERSIST
echo "size 1024 persistent rows 8"
TB_FORCE_ROWS=8 timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29539 tools/slab_bench.py --size 1024 2>&1 | tail -1
echo "size 1024 persistent rows 4"
TB_FORCE_ROWS=4 timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29539 tools/slab_bench.py --size 1024 2>&1 | tail -1
