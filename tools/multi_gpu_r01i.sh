# 8-GPU visit (round 1, final kernels): slab parity tests on 2 of the GPUs, chain-parallel bench on 8
timeout 300 python -m pytest tests/test_gpu_slab.py -x -q -m gpu 2>&1 | tail -3
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29538 bench.py --gpus 8 --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_r01i_8gpu.json
python -c "import json; d=json.load(open('gpurun_out/bench_r01i_8gpu.json')); print('bench gpus', d['n_gpus'], 'value', d['value'], 'e2e', d['e2e']['value'], 'hmc', d['hmc']['traj_per_sec']); print(json.dumps(d.get('other_configs'), indent=1)[:3000])"
