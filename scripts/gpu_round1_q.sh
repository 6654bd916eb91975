mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4)
timeout 900 python scripts/run_configs.py c3 --scale 20 --ef 1 2>&1 | tail -1 | cut -c1-700
timeout 900 python scripts/run_configs.py c3 --scale 22 --ef 1 --no-full-check 2>&1 | tail -1 | cut -c1-460
timeout 1200 python scripts/run_configs.py c3res --scale 22 --ef 4 2>&1 | tail -1 | cut -c1-600
