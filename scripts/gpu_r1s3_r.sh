mkdir -p gpurun_out
N=$1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 3 2>&1 | grep '"metric"' | tail -1 | tee gpurun_out/s3r_bench_n$N.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline']['kernel'], d['parity_spot_check'], d['e2e']['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"
