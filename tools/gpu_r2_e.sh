#!/bin/bash
# Round-2 GPU pass E (1 GPU): parity with the packed RRM kernel, RRM rates, RRM capture, per-component error survey.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
timeout 300 python benchmarks/n_sweep.py --models rrm-experimental --nsides 64,128,256,512,1024 > gpurun_out/r2e_rrm.jsonl 2> gpurun_out/r2e_rrm.err
ZODI_NO_X2=1 timeout 300 python benchmarks/n_sweep.py --models rrm-experimental --nsides 512 --label scalar_fused >> gpurun_out/r2e_rrm.jsonl 2>> gpurun_out/r2e_rrm.err
ZODI_FORCE_GENERIC=1 timeout 300 python benchmarks/n_sweep.py --models rrm-experimental --nsides 512 --label generic >> gpurun_out/r2e_rrm.jsonl 2>> gpurun_out/r2e_rrm.err
NCU_TAG=r2e bash benchmarks/ncu_round2_captures.sh rrm > gpurun_out/r2e_ncu.log 2>&1
timeout 300 python benchmarks/component_errors.py > gpurun_out/r2e_component_errors.jsonl 2> gpurun_out/r2e_component_errors.err
tail -3 gpurun_out/r2e_pytest.log; cut -c1-210 gpurun_out/r2e_rrm.jsonl; tail -1 gpurun_out/r2e_component_errors.jsonl
