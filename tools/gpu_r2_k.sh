#!/bin/bash
# Round-2 GPU pass K (1 GPU): final-build parity, A/B reference timings, ncu captures, launch list, bench lines.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log
AB_REPS=10 timeout 300 python benchmarks/ab_kernel.py > gpurun_out/r2k_ab.jsonl 2> gpurun_out/r2k_ab.err
timeout 400 python benchmarks/n_sweep.py --models planck18,dirbe --nsides 32,64,128,256,512,1024,2048 > gpurun_out/r2k_sweep.jsonl 2> gpurun_out/r2k_sweep.err
NCU_TAG=r2k bash benchmarks/ncu_round2_captures.sh x2a dirbe fp64a rrm > gpurun_out/r2k_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2k_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-configs > gpurun_out/r2k_bench_under_ncu.json 2> gpurun_out/r2k_bench_under_ncu.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo "bench rc=$?" >> gpurun_out/r2k_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2k_bench_ref.json 2>> gpurun_out/r2k_bench.err
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/r2k_pytest.log | tail -8
python - <<'PY'
import json
for l in open('gpurun_out/r2k_ab.jsonl'):
    d=json.loads(l); print(f"{d['case']:34s} min {d['ms_min']:.4f} med {d['ms_median']:.4f} chk {d['checksum']:.6f}")
d=json.loads(open('gpurun_out/r2k_bench.json').read())
print('bench value %.4e ms %.4f e2e %.3f fp64 %.3f'%(d['value'],d['ms_per_step'],d['e2e']['ms_per_step'],d['fp64_mode']['kernel_ms']))
for k,c in d['configs'].items(): print(k,{p:(round(c[p]['ms'],4),c[p]['ok']) for p in ('fp32','fp64')})
PY
tail -c 300 gpurun_out/r2k_bench.err
