# ncu --set full of the slab-ordered streaming SpMM (variants 3 and 0), one launch each
mkdir -p gpurun_out
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu"
SDB_SLAB_VARIANT=3 timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_stream -s 2 -c 1 -o gpurun_out/r1d_stream_v3 -f $B > gpurun_out/ncu_v3.log 2>&1
SDB_SLAB_VARIANT=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_stream -s 2 -c 1 -o gpurun_out/r1d_stream_v0 -f $B > gpurun_out/ncu_v0.log 2>&1
tail -3 gpurun_out/ncu_v3.log gpurun_out/ncu_v0.log
ls -la gpurun_out | tail -6
