mkdir -p gpurun_out
(timeout 500 python -m pytest tests/test_gpu_sharded.py -m gpu -q -x 2>&1 | tail -8) | tee gpurun_out/s3s_pytest.log
