mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
(timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5) | tee gpurun_out/s3a_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 | tee gpurun_out/s3a_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 2>&1 | tail -1 | tee gpurun_out/s3a_bench_ref.json
