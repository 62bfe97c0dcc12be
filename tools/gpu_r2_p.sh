#!/bin/bash
# Round-2 GPU pass P (8 GPUs): scaling curve with the persistent-tile kernel (A/B against ZODI_X2_PERSIST=0 at N=8).
mkdir -p gpurun_out
run() { local name=$1 n=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $n --steps 20 --warmup 5 "$@" > gpurun_out/r2p_$name.json 2> gpurun_out/r2p_$name.err; echo "$name rc=$?" >> gpurun_out/r2p_rc.log
}
run n8 8
ZODI_X2_PERSIST=0 run n8_plain 8 --no-e2e
run n4 4
run n2 2
timeout 600 python bench.py --steps 20 --warmup 5 --no-configs --no-cpu-baseline > gpurun_out/r2p_n1.json 2> gpurun_out/r2p_n1.err
run n8_b 8 --no-e2e
timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -q > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2p_pytest.log
cat gpurun_out/r2p_rc.log; tail -2 gpurun_out/r2p_pytest.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2p_n*.json')):
    try:
        d=json.loads(open(f).read()); sh=d.get('sharding') or {}
        print(f.split('r2p_')[1], '%.4e'%d['value'], 'ms %.4f kern %.4f'%(d['ms_per_step'],d['roofline']['kernel_ms']), 'rdv_us', sh.get('rendezvous_us'), 'ranks', [round(x,4) for x in sh.get('kernel_ms_per_rank',[])])
    except Exception as e: print(f,'ERR',e)
PY
