mkdir -p gpurun_out
(timeout 2400 python -m pytest tests -m gpu -q -x 2>&1 | tail -6) | tee gpurun_out/s3q_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 | tee gpurun_out/s3q_bench.json | cut -c1-400
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 2>&1 | tail -1 | tee gpurun_out/s3q_bench_ref.json | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6
