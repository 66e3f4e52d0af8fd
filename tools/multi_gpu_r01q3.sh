# 2-GPU visit: slab parity tests, then the one-launch slab solve at 2048^2, 1024^2, 512^2 (tools/slab_bench.py)
timeout 300 python -m pytest tests/test_gpu_slab.py -x -q -m gpu 2>&1 | tail -2
for size in 2048 1024 512; do
  echo "size $size persistent"
  timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29539 tools/slab_bench.py --size $size 2>&1 | tail -1
done
