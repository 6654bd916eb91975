mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -30) > gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
for cfg in "1 24" "0 24" "0 16" "0 32" "0 48" "0 12"; do
  set -- $cfg
  echo "SDB_SLAB=$1 SDB_SLAB_MB=$2"
  SDB_SLAB=$1 SDB_SLAB_MB=$2 timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['parity_spot_check'])"
done | tee gpurun_out/slab_sweep.log
