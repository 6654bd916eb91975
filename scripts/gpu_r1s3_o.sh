mkdir -p gpurun_out
N=$1
nvidia-smi -L | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 3 2>&1 | grep '"metric"' | tail -1 | tee gpurun_out/s3o_bench_n$N.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline']['kernel'], d['parity_spot_check'], d['e2e']['ms_per_step'], d['e2e']['value'], d['config']['parallelism'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus $N --steps 2 --warmup 3 2>&1 | grep '"impl"' | tail -1 | cut -c1-200
