mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv,noheader
run() { echo "== $*"; env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel'], d['clocks'])
except Exception as e: print('FAILED', e)
"; }
run SDB_SLAB=1
run SDB_SLAB=0
run SDB_SLAB=0
