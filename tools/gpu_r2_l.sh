#!/bin/bash
# Round-2 GPU pass L (1 GPU): parity after the chunk-independent launch shape, A/B of the cloud log-form,
# per-component errors.
mkdir -p gpurun_out; rm -f gpurun_out/r2l_ab.jsonl
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l_pytest.log
for rep in 1 2; do
for v in default cloud_direct; do
  if [ $v = default ]; then unset ZODI_B200_LIB; else export ZODI_B200_LIB=$PWD/zodipy_b200/build/variants/lib_$v.so; fi
  AB_REPS=10 timeout 300 python benchmarks/ab_kernel.py >> gpurun_out/r2l_ab.jsonl 2>> gpurun_out/r2l_ab.err
done
done
unset ZODI_B200_LIB
timeout 300 python benchmarks/component_errors.py > gpurun_out/r2l_component_errors.jsonl 2> gpurun_out/r2l_component_errors.err
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/r2l_pytest.log | tail -8
python - <<'PY'
import json
for l in open('gpurun_out/r2l_ab.jsonl'):
    d=json.loads(l); print(f"{d['lib'].split('/')[-1]:22s} {d['case']:34s} min {d['ms_min']:.4f} med {d['ms_median']:.4f} chk {d['checksum']:.6f}")
PY
tail -1 gpurun_out/r2l_component_errors.jsonl
