# usage: bash scripts/gpu_multi.sh N
N=$1
mkdir -p gpurun_out
nvidia-smi -L | head -8
nvidia-smi topo -m 2>/dev/null | head -12
for mode in fused nccl none; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --allgather $mode --no-cpu > gpurun_out/bench_n${N}_${mode}.log 2>&1
  tail -1 gpurun_out/bench_n${N}_${mode}.log | cut -c1-1500
done
