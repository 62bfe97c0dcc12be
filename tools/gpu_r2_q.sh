#!/bin/bash
# Round-2 GPU pass Q (1 GPU): final tree - full parity suite, smoke, bench line, multi-band occupancy A/B, ncu of the
# packed multi-band kernel.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2q_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2q_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2q_smoke.log
for v in default mb44 mb65; do
  if [ $v = default ]; then unset ZODI_B200_LIB; else export ZODI_B200_LIB=$PWD/zodipy_b200/build/variants/libzodi_$v.so; fi
  echo "{\"lib\": \"$v\"}" >> gpurun_out/r2q_multiband.jsonl
  timeout 200 python benchmarks/baseline_configs.py --multiband-only >> gpurun_out/r2q_multiband.jsonl 2>> gpurun_out/r2q_multiband.err
done
unset ZODI_B200_LIB
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; echo "bench rc=$?" >> gpurun_out/r2q_bench.err
NCU_TAG=r2q timeout 300 bash benchmarks/ncu_round2_captures.sh mbp > gpurun_out/r2q_ncu.log 2>&1
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/r2q_pytest.log | tail -8; tail -2 gpurun_out/r2q_smoke.log
python - <<'PY'
import json
for l in open('gpurun_out/r2q_multiband.jsonl'):
    d=json.loads(l)
    if 'lib' in d and len(d)==1: print('lib', d['lib']); continue
    print('  ', d['config'], 'packed %.4f scalar %.4f loop %.4f'%(d['multiband_ms'], d['scalar_multiband_ms'], d['per_band_loop_ms']), 'diff %.2e'%d['max_rel_diff_vs_single_band'])
d=json.loads(open('gpurun_out/r2q_bench.json').read())
print('bench value %.4e ms %.4f e2e %.3f fp64 %.3f'%(d['value'],d['ms_per_step'],d['e2e']['ms_per_step'],d['fp64_mode']['kernel_ms']))
for k,c in d['configs'].items(): print(k,{p:(round(c[p]['ms'],4),c[p]['ok']) for p in ('fp32','fp64')})
PY
tail -c 300 gpurun_out/r2q_bench.err gpurun_out/r2q_multiband.err; tail -3 gpurun_out/r2q_ncu.log
