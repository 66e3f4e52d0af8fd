timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cluster.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra --no-hmc | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value',d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'])"
timeout 600 python tools/probe_cluster.py 128,128,33,0.1 256,256,7,0.1 2>&1 | grep "solver=2" | cut -c1-330
