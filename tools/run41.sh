for n in 2 4 8; do
TB_SUBBATCHES=$n timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra --no-hmc | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('nsub=$n value',d['value'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])"
done
