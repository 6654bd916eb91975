mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_slab.py -m gpu -q -x 2>&1 | tail -3)
for cfg in "13 24" "16 24" "32 24"; do
  set -- $cfg
  echo "SDB_SLAB=2 SDB_SLAB_RPW=$1 SDB_SLAB_MB=$2"
  SDB_SLAB=2 SDB_SLAB_RPW=$1 SDB_SLAB_MB=$2 timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['parity_spot_check'])"
done | tee gpurun_out/slab_sweep4.log
