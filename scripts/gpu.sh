#!/bin/bash
# scripts/gpu.sh — the ONE runner for everything that needs the B200 box (replaces the one-shot
# scripts of round 1).  Run from the repo root, normally through gpurun:
#
#   gpurun --timeout 900 -- 'scripts/gpu.sh r2a tests bench launches "ncu:spmm_stream:stream" hashes'
#
# Every step logs to gpurun_out/<tag>_<step>.*; a failing step does not stop the following ones.
# Steps:
#   tests[=<pytest -k expression>]   pytest -m gpu (whole suite, or the selection)
#   testfile=<path>                  pytest -m gpu on one file
#   bench[=<extra bench.py args>]    python bench.py --steps 20 --warmup 5 <extra>   (1 GPU)
#   benchN=<n>[,<extra args>]        torchrun with n ranks
#   refarm                           bench.py --impl reference
#   launches[=<bench args>]          ncu launch list (gpu__time_duration.sum) of a short bench run
#   ncu=<kernel regex>,<name>[,<skip>[,<count>[,<python script + args>]]]   `--set full` capture of <count> launches
#   configs=<run_configs.py args>    scripts/run_configs.py ...
#   hashes                           SASS hashes of the profiled kernels (ties ncu bytes to this build)
#   py=<script and args>             python <script and args>
#   envpy=<label>,<VAR=VAL ..>,<script and args>   the same with environment variables set
#   sh=<label>,<shell command>       any shell command
#   smoke                            __graft_entry__.smoke()
set -u
cd "$(dirname "$0")/.."
TAG=$1; shift
OUT=gpurun_out
mkdir -p $OUT
NCU_COMMON="--clock-control none"
BENCH_SHORT="bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-legs --no-probes"
for step in "$@"; do
  name=${step%%=*}; arg=""; [[ "$step" == *=* ]] && arg=${step#*=}
  echo "=== [$TAG] $step ($(date +%T))"
  case $name in
    tests)
      if [ -n "$arg" ]; then timeout 1500 python -m pytest tests -x -q -m gpu -k "$arg" > $OUT/${TAG}_tests.log 2>&1
      else timeout 1500 python -m pytest tests -x -q -m gpu > $OUT/${TAG}_tests.log 2>&1; fi
      tail -5 $OUT/${TAG}_tests.log ;;
    testfile)
      timeout 1200 python -m pytest "$arg" -x -q -m gpu > $OUT/${TAG}_testfile_$(basename "$arg" .py).log 2>&1
      tail -5 $OUT/${TAG}_testfile_$(basename "$arg" .py).log ;;
    bench)
      timeout 1500 python bench.py --steps 20 --warmup 5 $arg > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
      tail -c 1500 $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err ;;
    benchN)
      n=${arg%%,*}; extra=""; [[ "$arg" == *,* ]] && extra=${arg#*,}
      timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
        --master-port 29511 bench.py --gpus $n --steps 20 --warmup 5 $extra \
        > $OUT/${TAG}_bench_n$n.json 2> $OUT/${TAG}_bench_n$n.err
      tail -c 2500 $OUT/${TAG}_bench_n$n.json; tail -5 $OUT/${TAG}_bench_n$n.err ;;
    refarm)
      timeout 900 python bench.py --impl reference --steps 5 --warmup 3 > $OUT/${TAG}_refarm.json 2> $OUT/${TAG}_refarm.err
      tail -c 600 $OUT/${TAG}_refarm.json ;;
    launches)
      timeout 900 ncu --metrics gpu__time_duration.sum $NCU_COMMON -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
        python ${arg:-$BENCH_SHORT} > $OUT/${TAG}_launches.log 2>&1
      tail -3 $OUT/${TAG}_launches.csv | cut -c1-300 ;;
    ncuenv)  # ncuenv=<VAR=VAL>,<kernel regex>,<name>,<skip>,<count>,<cmd>: the same with one environment variable set
      IFS=, read -r envs regex nm skip count cmd <<< "$arg"
      env $envs timeout 1200 ncu --set full $NCU_COMMON --import-source on -k "regex:$regex" -s ${skip:-2} -c ${count:-1} -f \
        -o $OUT/${TAG}_$nm python ${cmd:-$BENCH_SHORT} > $OUT/${TAG}_ncu_$nm.log 2>&1
      tail -2 $OUT/${TAG}_ncu_$nm.log
      python scripts/ncu_summary.py $OUT/${TAG}_$nm.ncu-rep > $OUT/${TAG}_${nm}_summary.txt 2>&1
      python scripts/ncu_traffic.py --dump $OUT/${TAG}_$nm.ncu-rep > $OUT/${TAG}_${nm}_launches.json 2>/dev/null
      if [ "$(stat -c %s $OUT/${TAG}_$nm.ncu-rep 2>/dev/null || echo 0)" -gt 8000000 ]; then rm -f $OUT/${TAG}_$nm.ncu-rep; fi ;;
    ncu)
      IFS=, read -r regex nm skip count cmd <<< "$arg"
      timeout 1200 ncu --set full $NCU_COMMON --import-source on -k "regex:$regex" -s ${skip:-2} -c ${count:-1} -f \
        -o $OUT/${TAG}_$nm python ${cmd:-$BENCH_SHORT} > $OUT/${TAG}_ncu_$nm.log 2>&1
      tail -2 $OUT/${TAG}_ncu_$nm.log
      # gpurun brings back at most 64 MiB: summarise on the box and keep the report itself only when it is small
      python scripts/ncu_summary.py $OUT/${TAG}_$nm.ncu-rep > $OUT/${TAG}_${nm}_summary.txt 2>&1
      python scripts/ncu_traffic.py --dump $OUT/${TAG}_$nm.ncu-rep > $OUT/${TAG}_${nm}_launches.json 2>/dev/null
      if [ "$(stat -c %s $OUT/${TAG}_$nm.ncu-rep 2>/dev/null || echo 0)" -gt 8000000 ]; then rm -f $OUT/${TAG}_$nm.ncu-rep; fi ;;
    configs)
      timeout 1500 python scripts/run_configs.py $arg > $OUT/${TAG}_configs.log 2>&1; cat $OUT/${TAG}_configs.log | cut -c1-1200 ;;
    hashes)
      python scripts/ncu_traffic.py --hashes > $OUT/${TAG}_sass_hashes.json 2>&1; cat $OUT/${TAG}_sass_hashes.json ;;
    py)
      timeout 1500 python $arg > $OUT/${TAG}_py_$(echo "$arg" | tr -c 'A-Za-z0-9' '_' | cut -c1-40).log 2>&1
      tail -40 $OUT/${TAG}_py_$(echo "$arg" | tr -c 'A-Za-z0-9' '_' | cut -c1-40).log ;;
    envpy)   # envpy=<label>,<VAR=VAL[ VAR=VAL..]>,<script and args>
      IFS=, read -r label envs cmd <<< "$arg"
      env $envs timeout 1500 python $cmd > $OUT/${TAG}_envpy_$label.log 2>&1; tail -25 $OUT/${TAG}_envpy_$label.log | cut -c1-1500 ;;
    sh)      # sh=<label>,<shell command>
      IFS=, read -r label cmd <<< "$arg"
      timeout 1500 bash -c "$cmd" > $OUT/${TAG}_sh_$label.log 2>&1; tail -30 $OUT/${TAG}_sh_$label.log | cut -c1-400 ;;
    smoke)
      timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; tail -8 $OUT/${TAG}_smoke.log ;;
    *) echo "unknown step $step" ;;
  esac
done
echo "=== [$TAG] done ($(date +%T))"
