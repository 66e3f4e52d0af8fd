for v in 1 0; do
echo "x_tmem=$v"
TB_RESIDENT_X_TMEM=$v python - <<'PY'
import sys, os, json
sys.path.insert(0, 'tools')
from probe import probe
for cfg in [(16,16,2048,0.1),(32,32,1024,0.1),(16,16,64,0.1),(32,32,64,0.1)]:
    o = probe(*cfg, solver=2, reps=3)
    print({k:o[k] for k in ('cfg','cg_ms','iters_mean')})
PY
done
