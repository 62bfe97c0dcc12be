#!/bin/bash
# Round-2 GPU pass M (8 GPUs): final scaling curve of the final build + one-process multi-device e2e.
mkdir -p gpurun_out
run() { local name=$1 n=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $n --steps 20 --warmup 5 "$@" > gpurun_out/r2m_$name.json 2> gpurun_out/r2m_$name.err; echo "$name rc=$?" >> gpurun_out/r2m_rc.log
}
timeout 600 python bench.py --steps 20 --warmup 5 --no-configs --no-cpu-baseline > gpurun_out/r2m_n1.json 2> gpurun_out/r2m_n1.err
run n2 2
run n4 4
run n8 8
run n8_nccl 8 --gather nccl --no-e2e
timeout 900 python benchmarks/multi_device_e2e.py > gpurun_out/r2m_multi_device_e2e.jsonl 2> gpurun_out/r2m_multi_device_e2e.err
timeout 900 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_parity.py -m gpu -q -k "fused or multi_device or two_processes" > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_pytest.log
cat gpurun_out/r2m_rc.log; tail -2 gpurun_out/r2m_pytest.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2m_n*.json')):
    try:
        d=json.loads(open(f).read()); sh=d['sharding']
        print(f.split('r2m_')[1], '%.4e'%d['value'], 'ms %.4f kern %.4f rdv_us %.1f'%(d['ms_per_step'],d['roofline']['kernel_ms'],sh['rendezvous_us']), 'e2e', d['e2e'] and (round(d['e2e']['ms_per_step'],3), round(d['e2e']['healpix_entry']['ms_per_step'],3), round(d['e2e']['lonlat_entry']['ms_per_step'],3)))
    except Exception as e: print(f,'ERR',e)
PY
cut -c1-190 gpurun_out/r2m_multi_device_e2e.jsonl
