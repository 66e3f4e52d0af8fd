timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -5
python - <<'PY'
import sys, json
sys.path.insert(0, '.'); sys.path.insert(0, 'tools')
from probe import probe
for cfg in [(256,256,64,0.01),(2048,2048,1,0.01),(128,128,2048,0.01),(64,64,256,0.1),(256,256,8,0.01)]:
    for solver in (1, 3):
        r = probe(*cfg, rows=0, solver=solver, reps=3, max_iter=151)
        print(cfg, 'solver', solver, 'us/iter', r['us_per_iter'], 'cg_frac(288B)', r['cg_frac'], flush=True)
PY
