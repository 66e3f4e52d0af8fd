timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5
for t in 0 28; do
TB_ROWS_PER_THREAD=$t timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra --no-hmc | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('tile=$t value',d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'])"
done
