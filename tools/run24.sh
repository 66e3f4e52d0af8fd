timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -4
for v in 1 0; do
TB_RESIDENT_X_TMEM=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra --no-hmc | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('tmem=$v value',d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'])"
done
