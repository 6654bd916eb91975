mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_slab.py -m gpu -q -x 2>&1 | tail -5) | tee gpurun_out/s3k_pytest.log
echo "== N=1"
timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel'], d['parity_spot_check'])"
for mode in fused none; do
  echo "== N=2 allgather=$mode"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --allgather $mode --no-e2e 2>&1 | grep '"metric"' | tail -1 | tee gpurun_out/s3k_bench_n2_$mode.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline']['kernel'], d['parity_spot_check'])"
done
