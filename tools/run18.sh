python - <<'PY'
import sys, json
sys.path.insert(0, '.'); sys.path.insert(0, 'tools')
from probe import probe
for cfg in [(32,32,1024,0.1),(32,32,592,0.1),(16,16,4096,0.1),(64,64,296,0.1),(64,64,148,0.1),(32,32,1,0.1),(64,64,1,0.1)]:
    for rows in (0, 44, 18):
        try:
            r = probe(*cfg, rows=rows, solver=2, reps=3)
            print(cfg, 'shape', rows, 'cg_ms', r['cg_ms'], 'iters', r['iters_max'], 'us/iter', r['us_per_iter'], 'site-iters/s %.3e' % (cfg[0]*cfg[1]*cfg[2]*r['iters_mean']/r['cg_ms']*1e3), flush=True)
        except Exception as e:
            print(cfg, rows, 'ERR', e)
PY
