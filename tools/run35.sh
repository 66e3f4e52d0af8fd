timeout 900 python -m pytest tests/test_gpu_cluster.py tests/test_gpu_family_b.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5
timeout 600 python bench.py --steps 10 --warmup 3 --no-hmc > gpurun_out/bench_r01j.json 2>gpurun_out/bench_r01j.err; tail -c 2500 gpurun_out/bench_r01j.json; tail -3 gpurun_out/bench_r01j.err
