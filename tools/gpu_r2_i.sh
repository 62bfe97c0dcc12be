#!/bin/bash
# Round-2 GPU pass I (8 GPUs): tail / imbalance variants of the fused gather, final scaling curve,
# one-process multi-device e2e with persistent workers + adaptive chunks.
mkdir -p gpurun_out
run() { # name, nproc, extra args..., env prefix via VAR
  local name=$1 n=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $n --steps 20 --warmup 5 "$@" > gpurun_out/r2i_$name.json 2> gpurun_out/r2i_$name.err; echo "$name rc=$?" >> gpurun_out/r2i_rc.log
}
run n8 8
ZODI_X2_THREADS=128 run n8_t128 8 --no-e2e
run n8_cb16k 8 --no-e2e --cyclic-block 16384
run n8_cb8k 8 --no-e2e --cyclic-block 8192
run n4 4
run n2 2
timeout 600 python bench.py --steps 20 --warmup 5 --no-configs --no-cpu-baseline > gpurun_out/r2i_n1.json 2> gpurun_out/r2i_n1.err
ZODI_X2_THREADS=64 timeout 600 python bench.py --steps 20 --warmup 5 --no-configs --no-cpu-baseline --no-e2e > gpurun_out/r2i_n1_t64.json 2> gpurun_out/r2i_n1_t64.err
timeout 900 python benchmarks/multi_device_e2e.py > gpurun_out/r2i_multi_device_e2e.jsonl 2> gpurun_out/r2i_multi_device_e2e.err
cat gpurun_out/r2i_rc.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2i_n*.json')):
    try:
        d=json.loads(open(f).read())
        sh=d['sharding']
        print(f.split('r2i_')[1], '%.4e'%d['value'], 'ms %.4f kern %.4f rdv_us %.1f'%(d['ms_per_step'],d['roofline']['kernel_ms'],sh['rendezvous_us']), 'ranks', [round(v,4) for v in sh['kernel_ms_per_rank']], sh['layout'])
    except Exception as e: print(f, 'ERR', e)
PY
cut -c1-190 gpurun_out/r2i_multi_device_e2e.jsonl
