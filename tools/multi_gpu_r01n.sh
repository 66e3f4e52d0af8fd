# 2-GPU visit (round 1, final tree): slab parity tests, chain-parallel bench and the reference arm under torchrun
timeout 300 python -m pytest tests/test_gpu_slab.py -x -q -m gpu 2>&1 | tail -3
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29538 bench.py --gpus 2 --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_r01n_2gpu.json
python -c "import json; d=json.load(open('gpurun_out/bench_r01n_2gpu.json')); print('bench gpus', d['n_gpus'], 'value', d['value'], 'e2e', d['e2e']['value'], 'hmc', d['hmc']['traj_per_sec']); print(json.dumps(d.get('other_configs'), indent=1)[:1500])"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29539 tools/slab_bench.py --size 2048 2>&1 | tail -2
