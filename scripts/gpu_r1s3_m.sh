run() { echo "== $*"; env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel'])
except Exception as e: print('FAILED', e)
"; }
run SDB200_LIB=$PWD/ab/libsdb200_old.so
run SDB_SLAB=0
