timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cg_gauge" 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra --no-hmc | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value',d['value'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])"
