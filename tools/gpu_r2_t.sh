#!/bin/bash
# Round-2 GPU pass T (1 GPU): final tree (ring / feature unrolled x5 at 64 registers) - full parity suite + smoke,
# then further occupancy / unroll variants over a few map sizes.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2t_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2t_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2t_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2t_smoke.log
for v in default rfu5c3 rfu10c4 thu2c4 thu2c5; do
  if [ $v = default ]; then unset ZODI_B200_LIB; else export ZODI_B200_LIB=$PWD/zodipy_b200/build/variants/libzodi_$v.so; fi
  timeout 100 python benchmarks/n_sweep.py --models dirbe --nsides 64,256,1024 --label $v >> gpurun_out/r2t_sweep.jsonl 2>> gpurun_out/r2t_sweep.err
done
for v in default thu2c4 thu2c5; do
  if [ $v = default ]; then unset ZODI_B200_LIB; else export ZODI_B200_LIB=$PWD/zodipy_b200/build/variants/libzodi_$v.so; fi
  timeout 100 python benchmarks/n_sweep.py --models planck18 --nsides 64,256,1024,2048 --label $v >> gpurun_out/r2t_sweep.jsonl 2>> gpurun_out/r2t_sweep.err
done
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/r2t_pytest.log | tail -8; tail -2 gpurun_out/r2t_smoke.log
python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/r2t_sweep.jsonl')]
for m in ('dirbe','planck18'):
    for ns in (64,256,1024,2048):
        r={x['lib']:round(x['ms'],5) for x in rows if x['nside']==ns and x['model']==m}
        cs={x['checksum'] for x in rows if x['nside']==ns and x['model']==m}
        if r: print(m, ns, r, 'same checksum' if len(cs)==1 else cs)
PY
tail -c 300 gpurun_out/r2t_sweep.err
