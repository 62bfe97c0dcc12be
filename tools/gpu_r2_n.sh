#!/bin/bash
# Round-2 GPU pass N (1 GPU): packed multi-band kernel (parity + timing), ring/feature unroll variants,
# launch shape on a rank's share of a split map.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multiband" > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n_pytest.log
timeout 600 python benchmarks/baseline_configs.py --multiband-only > gpurun_out/r2n_multiband.jsonl 2> gpurun_out/r2n_multiband.err
for v in default rfu2 rfu2c4 rfu5c4; do
  if [ $v = default ]; then unset ZODI_B200_LIB; else export ZODI_B200_LIB=$PWD/zodipy_b200/build/variants/libzodi_$v.so; fi
  AB_ONLY=dirbe AB_REPS=12 timeout 200 python benchmarks/ab_kernel.py >> gpurun_out/r2n_ab.jsonl 2>> gpurun_out/r2n_ab.err
done
unset ZODI_B200_LIB
timeout 300 python benchmarks/shard_shape.py > gpurun_out/r2n_shard_shape.jsonl 2> gpurun_out/r2n_shard_shape.err
tail -3 gpurun_out/r2n_pytest.log
cut -c1-600 gpurun_out/r2n_multiband.jsonl
python - <<'PY'
import json
for l in open('gpurun_out/r2n_ab.jsonl'):
    d=json.loads(l); print(f"{d['lib'][-24:]:26s}{d['case']:30s} min {d['ms_min']:.4f} med {d['ms_median']:.4f} chk {d['checksum']:.6f}")
for l in open('gpurun_out/r2n_shard_shape.jsonl'):
    d=json.loads(l); print(d['ranks'], d['n_los'], d['lanes'], round(d['ms'],4), '%.4e'%d['evals_per_s'])
PY
tail -c 400 gpurun_out/r2n_multiband.err gpurun_out/r2n_ab.err gpurun_out/r2n_shard_shape.err
