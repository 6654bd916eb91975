mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_slab.py -m gpu -q -x 2>&1 | tail -5) | tee gpurun_out/s3f_pytest.log
run() { echo "== $*"; env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['parity_spot_check'])
except Exception as e: print('FAILED', e)
"; }
(
for v in 0 1 2 3 4 5; do run SDB_SLAB_VARIANT=$v; done
for v in 0 2; do run SDB_SLAB_VARIANT=$v SDB_SLAB_MB=32; done
run SDB_SLAB_VARIANT=0 SDB_SLAB_MB=16
) 2>&1 | tee gpurun_out/s3f_sweep.log
