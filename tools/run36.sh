for r in 16 8 16 8; do
TB_FORCE_ROWS=$r timeout 600 python bench.py --steps 3 --warmup 3 --no-hmc --no-cpu-baseline --solver 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('rows=$r stream256', round(d['roofline_streaming']['us_per_iteration'],1), '2048:', round(d['other_configs']['2048x2048_single_lattice_1_gpu']['us_per_cg_iteration'],1), 'headline-stream us/iter', round(d['roofline']['us_per_iteration'],2))"
done
