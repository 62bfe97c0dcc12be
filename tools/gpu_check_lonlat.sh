#!/bin/bash
# One GPU-box session: new lon/lat tests, the default bench line, TOD end-to-end variants, parity suite.
mkdir -p gpurun_out
(timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_model_evaluate_standin.py -m gpu -x -q -k "lonlat or sky_rotation" > gpurun_out/r1c_new_tests.log 2>&1; echo "exit $?" >> gpurun_out/r1c_new_tests.log)
tail -5 gpurun_out/r1c_new_tests.log
(timeout 300 python bench.py > gpurun_out/r1c_bench_n1.json 2> gpurun_out/r1c_bench_n1.err; echo "bench exit $?")
tail -c 1500 gpurun_out/r1c_bench_n1.json
(timeout 150 python benchmarks/baseline_configs.py --tod-only > gpurun_out/r1c_tod_e2e.jsonl 2>&1; echo "tod exit $?")
tail -c 1200 gpurun_out/r1c_tod_e2e.jsonl
(timeout 330 python -m pytest tests/test_gpu_parity.py tests/test_model_evaluate_standin.py -m gpu -x -q > gpurun_out/r1c_gpu_parity.log 2>&1; echo "exit $?" >> gpurun_out/r1c_gpu_parity.log)
tail -5 gpurun_out/r1c_gpu_parity.log
