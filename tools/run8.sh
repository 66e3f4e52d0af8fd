python -m pytest tests -x -q -m gpu 2>&1 | tail -8
python bench.py --steps 10 --warmup 3 --no-cpu-baseline | tee gpurun_out/bench_r01d.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value',d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])"
for nsub in 1 2 8; do TB_SUBBATCHES=$nsub python bench.py --steps 10 --warmup 3 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('nsub $nsub value',d['value'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])"; done
