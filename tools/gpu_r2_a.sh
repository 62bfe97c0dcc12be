#!/bin/bash
# Round-2 GPU pass A: parity suite, packed-kernel shape sweep, A/B of kernel variants, short bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/r2a_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
timeout 600 python benchmarks/n_sweep.py --shapes --models planck18,dirbe --nsides 32,64,128,256,512,1024 > gpurun_out/r2a_sweep.jsonl 2> gpurun_out/r2a_sweep.err
timeout 200 python benchmarks/n_sweep.py --models planck18,dirbe --nsides 2048 >> gpurun_out/r2a_sweep.jsonl 2>> gpurun_out/r2a_sweep.err
for v in default rcpnr rf4; do
  if [ $v = default ]; then unset ZODI_B200_LIB; else export ZODI_B200_LIB=$PWD/zodipy_b200/build/variants/lib_$v.so; fi
  AB_REPS=10 timeout 300 python benchmarks/ab_kernel.py >> gpurun_out/r2a_ab.jsonl 2>> gpurun_out/r2a_ab.err
done
unset ZODI_B200_LIB
ZODI_NO_RING_TPOLY=1 AB_REPS=10 AB_ONLY=nside1024 timeout 300 python benchmarks/ab_kernel.py | sed 's/"lib": "default"/"lib": "no_ring_tpoly"/' >> gpurun_out/r2a_ab.jsonl 2>> gpurun_out/r2a_ab.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -3 gpurun_out/r2a_pytest.log
