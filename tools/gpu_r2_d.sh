#!/bin/bash
# Round-2 GPU pass D (1 GPU): parity, RRM rates after the MUFU trims, final-build ncu captures, bench line.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
timeout 300 python benchmarks/n_sweep.py --models rrm-experimental --nsides 64,256,512 > gpurun_out/r2d_rrm.jsonl 2> gpurun_out/r2d_rrm.err
timeout 300 python benchmarks/n_sweep.py --models rrm-experimental --nsides 256 --precision fp64 >> gpurun_out/r2d_rrm.jsonl 2>> gpurun_out/r2d_rrm.err
NCU_TAG=r2d bash benchmarks/ncu_round2_captures.sh x2a dirbe fp64a rrm > gpurun_out/r2d_ncu.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench rc=$?" >> gpurun_out/r2d_bench.err
tail -3 gpurun_out/r2d_pytest.log; tail -c 300 gpurun_out/r2d_bench.err; cat gpurun_out/r2d_rrm.jsonl | cut -c1-200
