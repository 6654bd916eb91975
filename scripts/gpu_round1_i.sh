mkdir -p gpurun_out
for mode in 1 2; do
  echo "SDB_DENSE_MODE=$mode"
  SDB_DENSE_MODE=$mode timeout 900 python scripts/run_configs.py c4 2>&1 | tail -1 | cut -c1-330
  SDB_DENSE_MODE=$mode timeout 900 python scripts/run_configs.py c4 --gram-m 200000 --gram-n 10000 2>&1 | tail -1 | cut -c1-330
done | tee gpurun_out/dense_modes.log
for st in 2 3 4; do
  echo "SDB_BSR_STAGES=$st"
  SDB_BSR_STAGES=$st timeout 600 python scripts/run_configs.py c5bsr 2>&1 | tail -1 | cut -c1-330
done | tee gpurun_out/bsr_stages.log
SDB_DENSE_MODE=2 timeout 1200 python -m pytest tests -m gpu -q -x -k "gram or dense" 2>&1 | tail -3
