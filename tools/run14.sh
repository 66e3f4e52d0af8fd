python - <<'PY'
import sys, json
sys.path.insert(0, '.')
sys.path.insert(0, 'tools')
from probe import probe
for cfg in [(256,256,64,0.01),(2048,2048,1,0.01),(64,64,1024,0.01),(128,128,512,0.01)]:
    for rows in (2,4,8,16):
        r = probe(*cfg, rows=rows, solver=1, reps=5, max_iter=101)
        print(cfg, 'rows', rows, 'apply_us', r['apply_us'], 'apply_frac', r['apply_frac'], 'us/iter', r['us_per_iter'], 'cg_frac', r['cg_frac'], flush=True)
PY
