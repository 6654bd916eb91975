mkdir -p gpurun_out
N=$1
echo skip-tests
for strat in ce stores; do
  echo "== N=$N fused, SDB_ALLGATHER=$strat"
  SDB_ALLGATHER=$strat timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --allgather fused --no-e2e 2>&1 | grep '"metric"' | tail -1 | tee gpurun_out/s3p_bench_n${N}_$strat.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline']['kernel'], d['parity_spot_check'], d['gpu_launches'])"
done
