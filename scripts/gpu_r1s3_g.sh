mkdir -p gpurun_out
(timeout 2400 python -m pytest tests -m gpu -q -x 2>&1 | tail -8) | tee gpurun_out/s3g_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 | tee gpurun_out/s3g_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 2>&1 | tail -1 | tee gpurun_out/s3g_bench_ref.json
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1d_launches_bench.csv $B > gpurun_out/ncu_g1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_stream -s 2 -c 1 -o gpurun_out/r1d_stream -f $B > gpurun_out/ncu_g2.log 2>&1
echo "== c5 shard (N=256): K1 vs stream"
SDB_SLAB=1 timeout 600 python scripts/run_configs.py c5 2>&1 | tail -1 | cut -c1-400
SDB_SLAB=2 timeout 600 python scripts/run_configs.py c5 2>&1 | tail -1 | cut -c1-400
ls -la gpurun_out | tail -8
