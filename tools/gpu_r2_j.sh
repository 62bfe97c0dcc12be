#!/bin/bash
# Round-2 GPU pass J (1 GPU): A/B of the band-pretest (MUFU floor) and ring/feature-first variants.
mkdir -p gpurun_out; rm -f gpurun_out/r2j_ab.jsonl
for rep in 1 2; do
for v in default pretest rffirst both; do
  if [ $v = default ]; then unset ZODI_B200_LIB; else export ZODI_B200_LIB=$PWD/zodipy_b200/build/variants/lib_$v.so; fi
  AB_REPS=10 timeout 300 python benchmarks/ab_kernel.py >> gpurun_out/r2j_ab.jsonl 2>> gpurun_out/r2j_ab.err
done
done
unset ZODI_B200_LIB
python - <<'PY'
import json
for l in open('gpurun_out/r2j_ab.jsonl'):
    d=json.loads(l); print(f"{d['lib'].split('/')[-1]:18s} {d['case']:34s} min {d['ms_min']:.4f} med {d['ms_median']:.4f} chk {d['checksum']:.6f}")
PY
