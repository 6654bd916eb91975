mkdir -p gpurun_out
(timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_reference_live.py -m gpu -q -x -k "spgemm or gram or sparse_sparse or syrk" 2>&1 | tail -5) | tee gpurun_out/s3i_pytest.log
SDB_TRACE=1 timeout 900 python scripts/run_configs.py c3 --scale 22 --ef 1 --no-full-check 2>&1 | grep -E "spgemm|config" | tail -40 | cut -c1-700 | tee gpurun_out/s3i_c3.log
