timeout 600 python -m pytest tests/test_gpu_slab.py -x -q -m gpu 2>&1 | tail -8
timeout 300 python tools/slab_bench.py
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/slab_bench.py 2>&1 | tail -1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 tools/slab_bench.py --size 4096 2>&1 | tail -1
timeout 300 python tools/slab_bench.py --size 4096
