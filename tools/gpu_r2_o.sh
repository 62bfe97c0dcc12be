#!/bin/bash
# Round-2 GPU pass O (2 GPUs): persistent-tile packed kernel with peer stores - parity and A/B at N=2.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_parity.py -m gpu -q -k "persistent or fused or multiband_packed" > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2o_pytest.log
run() { local name=$1; shift
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e "$@" > gpurun_out/r2o_$name.json 2> gpurun_out/r2o_$name.err; echo "$name rc=$?" >> gpurun_out/r2o_rc.log
}
ZODI_X2_PERSIST=0 run n2_plain
run n2_persist
ZODI_X2_PERSIST=0 run n2_plain_b
run n2_persist_b
for p in 0 2; do ZODI_X2_PERSIST=$p AB_ONLY=fp32 AB_REPS=10 timeout 200 python benchmarks/ab_kernel.py > gpurun_out/r2o_ab_persist$p.jsonl 2>> gpurun_out/r2o_ab.err; done
tail -4 gpurun_out/r2o_pytest.log; cat gpurun_out/r2o_rc.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2o_n2*.json')):
    try:
        d=json.loads(open(f).read()); sh=d['sharding']
        print(f.split('r2o_')[1], '%.4e'%d['value'], 'ms %.4f'%d['ms_per_step'], 'kern', [round(x,4) for x in sh['kernel_ms_per_rank']], 'rdv_us %.1f'%sh['rendezvous_us'])
    except Exception as e: print(f,'ERR',e)
for p in (0,2):
    for l in open(f'gpurun_out/r2o_ab_persist{p}.jsonl'):
        d=json.loads(l); print('persist',p,f"{d['case']:34s} min {d['ms_min']:.4f} med {d['ms_median']:.4f} chk {d['checksum']:.6f}")
PY
tail -c 600 gpurun_out/r2o_n2_persist.err gpurun_out/r2o_ab.err
