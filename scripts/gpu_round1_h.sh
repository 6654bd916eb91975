mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -30) > gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
SDB_TRACE=1 timeout 900 python scripts/run_configs.py c4 > gpurun_out/configs_c4_full.log 2>&1; tail -12 gpurun_out/configs_c4_full.log
timeout 900 python scripts/run_configs.py c3 --scale 20 --ef 1 > gpurun_out/configs_c3_s20.log 2>&1; tail -1 gpurun_out/configs_c3_s20.log
