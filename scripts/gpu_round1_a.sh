mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -80) > gpurun_out/pytest_gpu.log
tail -60 gpurun_out/pytest_gpu.log
# launch list of the bench command (cold-cache, serialised: compare shares)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
tail -3 gpurun_out/launches_r1.csv
# full capture of the SpMM kernel
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_rowmajor -s 3 -c 2 -o gpurun_out/spmm_r1 -f python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
