mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -30) > gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 900 python scripts/run_configs.py c3 --scale 20 --ef 1 > gpurun_out/configs_c3_s20.log 2>&1; tail -1 gpurun_out/configs_c3_s20.log
timeout 900 python scripts/run_configs.py c3 --scale 22 --ef 1 --no-full-check > gpurun_out/configs_c3_s22.log 2>&1; tail -1 gpurun_out/configs_c3_s22.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log | cut -c1-400
