#!/bin/bash
# Round-2 GPU pass F (8 GPUs): scaling of the fused gather with the flag rendezvous, NCCL variant,
# one-process multi-device e2e, multi-rank tests.
mkdir -p gpurun_out
for n in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2f_bench_n$n.json 2> gpurun_out/r2f_bench_n$n.err; echo "bench n=$n rc=$?" >> gpurun_out/r2f_bench_n$n.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --steps 20 --warmup 5 --gather nccl --no-e2e > gpurun_out/r2f_bench_n8_nccl.json 2> gpurun_out/r2f_bench_n8_nccl.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 --shard contiguous --no-e2e > gpurun_out/r2f_bench_n8_contig.json 2> gpurun_out/r2f_bench_n8_contig.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-configs --no-cpu-baseline > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err
timeout 900 python benchmarks/multi_device_e2e.py > gpurun_out/r2f_multi_device_e2e.jsonl 2> gpurun_out/r2f_multi_device_e2e.err
timeout 900 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_parity.py -m gpu -x -q -k "fused or multi_device or two_processes" > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
tail -3 gpurun_out/r2f_pytest.log; for n in 1 2 4 8; do python -c "
import json,sys
d=json.loads(open('gpurun_out/r2f_bench_n$n.json').read()); print($n, '%.4e'%d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e'] and d['e2e']['ms_per_step'])"; done; cut -c1-200 gpurun_out/r2f_multi_device_e2e.jsonl
