mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -40) > gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
timeout 600 python scripts/run_configs.py c1 c5bsr > gpurun_out/configs_c.log 2>&1; tail -3 gpurun_out/configs_c.log
timeout 900 python scripts/run_configs.py c4 > gpurun_out/configs_c4_full.log 2>&1; tail -3 gpurun_out/configs_c4_full.log
timeout 900 python scripts/run_configs.py c3 --scale 22 --ef 1 --no-full-check > gpurun_out/configs_c3_s22.log 2>&1; tail -3 gpurun_out/configs_c3_s22.log
