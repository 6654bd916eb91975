mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40) > gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2>&1; tail -2 gpurun_out/bench.log
timeout 900 python scripts/run_configs.py c1 c3 --scale 20 --ef 1 > gpurun_out/configs_a.log 2>&1; tail -5 gpurun_out/configs_a.log
timeout 900 python scripts/run_configs.py c4 --gram-m 200000 --gram-n 10000 > gpurun_out/configs_b.log 2>&1; tail -3 gpurun_out/configs_b.log
timeout 900 python scripts/run_configs.py c5bsr > gpurun_out/configs_c.log 2>&1; tail -3 gpurun_out/configs_c.log
