mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_slab.py -m gpu -q -x 2>&1 | tail -5) | tee gpurun_out/s3h_pytest.log
for mode in fused nccl none; do
  echo "== allgather=$mode"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --allgather $mode 2>&1 | grep '"metric"' | tail -1 | tee gpurun_out/s3h_bench_n2_$mode.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline']['kernel'], d['parity_spot_check'], d['e2e']['ms_per_step'] if d.get('e2e') else None)"
done
