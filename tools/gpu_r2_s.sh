#!/bin/bash
# Round-2 GPU pass S (1 GPU): ring / feature loop unroll + 64 registers over the whole size range (dirbe).
mkdir -p gpurun_out
for v in default rfu5c4 rfu2c4; do
  if [ $v = default ]; then unset ZODI_B200_LIB; else export ZODI_B200_LIB=$PWD/zodipy_b200/build/variants/libzodi_$v.so; fi
  timeout 200 python benchmarks/n_sweep.py --models dirbe --nsides 32,64,128,256,512,1024 --label $v >> gpurun_out/r2s_sweep.jsonl 2>> gpurun_out/r2s_sweep.err
done
python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/r2s_sweep.jsonl')]
for ns in (32,64,128,256,512,1024):
    print(ns, {r['lib']:(round(r['ms'],5), r['lanes']) for r in rows if r['nside']==ns}, {r['lib']:r['checksum'] for r in rows if r['nside']==ns and r['lib']!='default'} )
PY
tail -c 300 gpurun_out/r2s_sweep.err
