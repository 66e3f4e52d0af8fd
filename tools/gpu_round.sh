#!/usr/bin/env bash
# One GPU visit: parity tests, smoke, bench, ncu launch list + full capture of the top kernels.
# Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh [tag]
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $OUT/smi_$TAG.txt 2>&1
echo "== pytest -m gpu"; python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee $OUT/pytest_gpu_$TAG.log
echo "== smoke"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke_$TAG.log
echo "== bench"; python bench.py --steps 10 --warmup 3 2>&1 | tail -2 | tee $OUT/bench_$TAG.json
echo "== bench reference arm"; python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee $OUT/bench_ref_$TAG.json
echo "== bench streaming solver"; python bench.py --steps 5 --warmup 3 --solver 1 --no-cpu-baseline --no-hmc --no-extra 2>&1 | tail -1 | tee $OUT/bench_stream_$TAG.json
echo "== ncu launch list (bench command, resident solver)"
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-hmc --no-extra > $OUT/ncu_launch_$TAG.log 2>&1
tail -1 $OUT/ncu_launch_$TAG.log | cut -c1-200
echo "== ncu full: resident kernel"
ncu --set full --clock-control none --import-source on -k regex:resident -s 2 -c 1 \
    -f -o $OUT/prof_${TAG}_resident python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-hmc --no-extra > $OUT/ncu_full_$TAG.log 2>&1
tail -1 $OUT/ncu_full_$TAG.log | cut -c1-200
echo "== ncu full: cluster kernel (128x128 x 33 chains = one wave of 4-CTA clusters)"
ncu --set full --clock-control none --import-source on -k regex:cluster_cg -s 1 -c 1 \
    -f -o $OUT/prof_${TAG}_cluster python tools/probe_cluster.py 128,128,33,0.1 > $OUT/ncu_full_cluster_$TAG.log 2>&1
tail -1 $OUT/ncu_full_cluster_$TAG.log | cut -c1-200
echo "== ncu full: planned cluster launch (256x256 x 8 chains on 7 co-resident clusters of 16 CTAs)"
ncu --set full --clock-control none --import-source on -k regex:cluster_cg -s 1 -c 1 \
    -f -o $OUT/prof_${TAG}_cluster_planned python tools/probe_cluster.py 256,256,8,0.05 > $OUT/ncu_full_cluster_planned_$TAG.log 2>&1
tail -1 $OUT/ncu_full_cluster_planned_$TAG.log | cut -c1-200
echo "== ncu full: TMA-staged streaming kernels (2048x2048 single lattice, 1.3 GB working set)"
ncu --set full --clock-control none --import-source on -k regex:'dslash_pipe|xpay' -s 30 -c 6 \
    -f -o $OUT/prof_${TAG}_staged python tools/probe.py --one 2048 2048 1 > $OUT/ncu_full_staged_$TAG.log 2>&1
tail -1 $OUT/ncu_full_staged_$TAG.log | cut -c1-200
echo "== ncu full: streaming kernels on a working set > L2 (256x256 x 64 chains)"
ncu --set full --clock-control none --import-source on -k regex:'dslash_kernel|axpy_norm|xpay' -s 40 -c 8 \
    -f -o $OUT/prof_${TAG}_stream python tools/probe.py --one 256 256 64 > $OUT/ncu_full_stream_$TAG.log 2>&1
tail -1 $OUT/ncu_full_stream_$TAG.log | cut -c1-200
echo "== summaries of the captures (gpurun brings back at most 64 MiB: only the report of the bench's top kernel travels)"
for k in resident cluster cluster_planned staged stream; do
  python tools/ncu_summary.py full $OUT/prof_${TAG}_$k.ncu-rep > $OUT/ncu_full_${TAG}_$k.txt 2>&1
  [ $k = resident ] || rm -f $OUT/prof_${TAG}_$k.ncu-rep
done
python tools/ncu_summary.py launches $OUT/launches_$TAG.csv > $OUT/launches_${TAG}_bench.txt 2>&1
echo "== planned vs plain launches, staged vs marching kernels"
python tools/probe_plan.py 2>&1 | tee $OUT/probe_plan_$TAG.txt | cut -c1-250
python tools/probe_pipe.py 2048,2048,1,0.01 4096,4096,1,0.01 512,512,16,0.01 256,256,64,0.01 128,128,2048,0.01 2>&1 | tee $OUT/probe_pipe_$TAG.txt | cut -c1-250
echo "== many sources on one gauge field (tb_set_gauge_shared), cluster solver"
python tools/probe_shared.py 2>&1 | tee $OUT/probe_shared_$TAG.txt | cut -c1-250
python tools/probe_cluster.py 256,256,8,0.01 256,256,7,0.05 128,128,33,0.1 2>&1 | grep "solver=2" | tee $OUT/probe_cluster_$TAG.txt | cut -c1-250
echo "== compute-sanitizer over the round-2 kernels"
for tool in memcheck racecheck synccheck; do
  compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize.py r2 2>&1 | tail -3 | tee $OUT/sanitize_${TAG}_$tool.txt
done
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  echo "== 2+ GPUs: slab parity tests and the one-launch slab solve (tools/multi_gpu.sh)"
  bash tools/multi_gpu.sh $TAG quick 2>&1 | tee $OUT/slab_$TAG.txt | cut -c1-250
fi
ls -la $OUT | tail -15
