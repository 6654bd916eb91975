mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5)
timeout 600 python scripts/run_configs.py c5 c5bsr c1 2>&1 | tail -3 | cut -c1-400
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log | cut -c1-300
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -6
