mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -30) > gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 900 python scripts/run_configs.py c3 --scale 20 --ef 1 > gpurun_out/configs_c3_s20.log 2>&1; tail -1 gpurun_out/configs_c3_s20.log
timeout 900 python scripts/run_configs.py c3 --scale 22 --ef 1 --no-full-check > gpurun_out/configs_c3_s22.log 2>&1; tail -1 gpurun_out/configs_c3_s22.log
for t in 5 6 7 8 9 10 2; do
  echo "SDB_SPMM_TUNE=$t"
  SDB_SPMM_TUNE=$t timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'])"
done | tee gpurun_out/spmm_tune2.log
