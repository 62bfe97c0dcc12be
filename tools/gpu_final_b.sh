#!/bin/bash
# GPU box session B: the default bench line, then the ncu launch list of a short bench run.
mkdir -p gpurun_out
(timeout 300 python bench.py > gpurun_out/r1b_bench_n1.json 2> gpurun_out/r1b_bench_n1.err; echo "bench exit $?")
tail -c 400 gpurun_out/r1b_bench_n1.json
(timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1b_launches_bench_steps3.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r1b_bench_under_ncu.log 2>&1; echo "ncu exit $?")
wc -l gpurun_out/r1b_launches_bench_steps3.csv
