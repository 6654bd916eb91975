for kb in 32 64 96; do echo "== XKB $kb"; SDB_SPMV_TILE_XKB=$kb python scripts/run_configs.py spmv spmv64 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['dtype'], 'tile_ms', round(d['tile_ms'],4), 'first', round(d['tile_first_call_ms'],2), 'wide', round(d['wide_ms'],4), 'err', d['tile_max_rel_err'])
"; done
