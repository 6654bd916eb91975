# ncu evidence for round 1 (one GPU).  Numbers under the profiler are never bench values.
mkdir -p gpurun_out
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1b_launches_bench.csv $B > gpurun_out/ncu_a.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_rowmajor -s 3 -c 1 -o gpurun_out/r1b_spmm -f $B > gpurun_out/ncu_b.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_bsr -s 2 -c 1 -o gpurun_out/r1b_bsr -f python scripts/run_configs.py c5bsr > gpurun_out/ncu_c.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spgemm_ -s 6 -c 6 -o gpurun_out/r1b_spgemm -f python scripts/run_configs.py c3 --scale 19 --ef 1 --no-full-check > gpurun_out/ncu_d.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spgemm_dense -s 1 -c 1 -o gpurun_out/r1b_syrkd -f python scripts/run_configs.py c4 --gram-m 400000 --gram-n 20000 > gpurun_out/ncu_e.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1b_launches_spgemm.csv python scripts/run_configs.py c3 --scale 19 --ef 1 --no-full-check > gpurun_out/ncu_f.log 2>&1
ls -la gpurun_out | tail -12
