timeout 600 python -m pytest tests/test_gpu_slab.py -x -q -m gpu 2>&1 | tail -15
timeout 600 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_slab.py 2>&1 | tail -5
