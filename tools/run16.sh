timeout 900 python -m pytest tests/test_gpu_family_b.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -25
