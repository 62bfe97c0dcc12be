#!/bin/bash
# Round-2 GPU pass C (2 GPUs): multi-rank / multi-device tests, 2-rank bench (fused gather with the flag
# rendezvous), one-process multi-device e2e.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2c_bench_n2.json 2> gpurun_out/r2c_bench_n2.err; echo "bench rc=$?" >> gpurun_out/r2c_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --gather nccl --no-e2e > gpurun_out/r2c_bench_n2_nccl.json 2>> gpurun_out/r2c_bench_n2.err
timeout 600 python benchmarks/multi_device_e2e.py > gpurun_out/r2c_multi_device_e2e.jsonl 2> gpurun_out/r2c_multi_device_e2e.err
tail -4 gpurun_out/r2c_pytest.log; tail -c 400 gpurun_out/r2c_bench_n2.err; cat gpurun_out/r2c_multi_device_e2e.jsonl | cut -c1-220
