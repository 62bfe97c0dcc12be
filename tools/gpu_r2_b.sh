#!/bin/bash
# Round-2 GPU pass B: parity (fused RRM kernel is new), ncu source-level captures, RRM rates, full bench line.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
timeout 300 python benchmarks/n_sweep.py --models rrm-experimental --nsides 64,256,512 > gpurun_out/r2b_rrm.jsonl 2> gpurun_out/r2b_rrm.err
ZODI_FORCE_GENERIC=1 timeout 300 python benchmarks/n_sweep.py --models rrm-experimental --nsides 512 --label generic >> gpurun_out/r2b_rrm.jsonl 2>> gpurun_out/r2b_rrm.err
timeout 300 python benchmarks/n_sweep.py --models rrm-experimental --nsides 256 --precision fp64 >> gpurun_out/r2b_rrm.jsonl 2>> gpurun_out/r2b_rrm.err
ZODI_FORCE_GENERIC=1 timeout 300 python benchmarks/n_sweep.py --models rrm-experimental --nsides 256 --precision fp64 --label generic >> gpurun_out/r2b_rrm.jsonl 2>> gpurun_out/r2b_rrm.err
bash benchmarks/ncu_round2_captures.sh dirbe x2 rrm > gpurun_out/r2b_ncu.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?" >> gpurun_out/r2b_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2b_bench_ref.json 2>> gpurun_out/r2b_bench.err
tail -3 gpurun_out/r2b_pytest.log
tail -c 600 gpurun_out/r2b_bench.err
ls -la gpurun_out | head -40
