ncu --set full --clock-control none --import-source on -k regex:resident_wt -s 2 -c 1 \
    -f -o gpurun_out/prof_r01g_resident_wt python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-hmc --no-extra > gpurun_out/ncu_full_r01g.log 2>&1
tail -2 gpurun_out/ncu_full_r01g.log | cut -c1-300
