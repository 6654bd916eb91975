N=$1
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_n${N}_fused_b.log 2>&1
tail -1 gpurun_out/bench_n${N}_fused_b.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['n_gpus'], 'ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e'], d['parity_spot_check'])"
nvidia-smi topo -m | head -12
