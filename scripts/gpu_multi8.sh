N=$1
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for mode in fused nccl; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --allgather $mode --no-cpu > gpurun_out/bench_n${N}_${mode}.log 2>&1
  tail -1 gpurun_out/bench_n${N}_${mode}.log | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('$mode', d['n_gpus'], 'ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['ms_per_step'] if d.get('e2e') else None, d['parity_spot_check'])
except Exception as e: print('PARSE FAIL', e)
"
  tail -3 gpurun_out/bench_n${N}_${mode}.log | cut -c1-300 | grep -v '^{' 
done
