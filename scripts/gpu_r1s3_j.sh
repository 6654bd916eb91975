mkdir -p gpurun_out
for mb in 2048 192 128 96 64 48; do
  echo "== SDB_WIDE_SCRATCH_MB=$mb"
  SDB_WIDE_SCRATCH_MB=$mb SDB_TRACE=1 timeout 900 python scripts/run_configs.py c3 --scale 22 --ef 1 --no-full-check 2>&1 | grep -E "wide bin|\"config\"" | tail -3 | cut -c1-330
done 2>&1 | tee gpurun_out/s3j_wide_scratch.log
