# 8-GPU visit: chain-parallel bench scaling + slab bench
for n in 8 4; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 10 --warmup 3 --no-extra 2>&1 | tail -1 > gpurun_out/bench_r01f_${n}gpu.json
python -c "import json; d=json.load(open('gpurun_out/bench_r01f_${n}gpu.json')); print('bench gpus', d['n_gpus'], 'value', d['value'], 'e2e', d['e2e']['value'], 'hmc', d['hmc']['traj_per_sec'])"
done
for n in 8 4; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n tools/slab_bench.py 2>&1 | tail -1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2955$n tools/slab_bench.py --size 4096 2>&1 | tail -1
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 tools/slab_bench.py --size 8192 --iters 100 2>&1 | tail -1
