# First GPU call of the next round (ideas from DESIGN.md §7, none measured yet).  Usage:
#   gpurun --timeout 1500 -- 'bash scripts/gpu_next_round.sh 1'      # regression + variant check on one GPU
#   gpurun --gpus 8 --timeout 900 -- 'bash scripts/gpu_next_round.sh 8'   # the N = 8 exchange (8x the GPU-minutes!)
N=${1:-1}
mkdir -p gpurun_out
line() { python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline']['kernel'], d['parity_spot_check'])"; }
if [ "$N" = "1" ]; then
  (timeout 2400 python -m pytest tests -m gpu -q -x 2>&1 | tail -4) | tee gpurun_out/next_pytest.log
  timeout 600 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 | tee gpurun_out/next_bench.json | cut -c1-300
  # is the 1-CTA x 32-warp shape (416 rows / SM) only slow because ptxas pairs its gather ring?  (scripts/sass_gaps.py)
  for v in 0 5; do echo "== SDB_SLAB_VARIANT=$v"; SDB_SLAB_VARIANT=$v timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu 2>&1 | tail -1 | line; done
else
  # the exchange at N ranks: automatic choice vs every explicit strategy (K1s + stores at N = 8 has never been measured)
  for strat in auto k1 stores ce; do
    echo "== N=$N SDB_ALLGATHER=$strat"
    SDB_ALLGATHER=$strat timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 20 --warmup 3 --allgather fused --no-e2e 2>&1 | grep '"metric"' | tail -1 | tee gpurun_out/next_bench_n${N}_$strat.json | line
  done
fi
