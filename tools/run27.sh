timeout 900 python -m pytest tests/test_gpu_cluster.py -x -q -m gpu 2>&1 | tail -15
python - <<'PY'
import thirring2d_b200 as tb
for nt,nx in [(64,64),(128,128),(256,256),(128,64),(256,128),(128,256),(64,256)]:
    with tb.Context(nt,nx,2,tb.MODE_ADJOINT) as c: print(nt,nx,c.solver_info())
PY
timeout 600 python tools/probe_cluster.py 2>&1 | tail -14
