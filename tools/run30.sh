timeout 900 python -m pytest tests/test_gpu_cluster.py -x -q -m gpu 2>&1 | tail -8
timeout 600 python tools/probe_cluster.py 2>&1 | grep "solver=2" | cut -c1-330
