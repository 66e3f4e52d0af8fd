python -m pytest tests -x -q -m gpu 2>&1 | tail -8
for shape in 44 18 28 24; do echo "shape $shape"; python bench.py --steps 10 --warmup 3 --no-cpu-baseline --rows $shape | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['us_per_iteration'], d['e2e']['value'])"; done
