mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5)
for i in 1 2; do timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'])"; done
timeout 600 python scripts/run_configs.py c5 2>&1 | tail -1 | cut -c1-300
