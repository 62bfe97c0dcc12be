#!/bin/bash
# Two-GPU box session: multi-device host path, multi-rank tests, 2-rank bench line.
mkdir -p gpurun_out
(timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multirank.py -m gpu -x -q -k "multi_device or multirank or two_processes or fused_peer" > gpurun_out/r1c_2gpu_tests.log 2>&1; echo "exit $?" >> gpurun_out/r1c_2gpu_tests.log)
tail -5 gpurun_out/r1c_2gpu_tests.log
(timeout 200 python benchmarks/multi_device_e2e.py > gpurun_out/r1c_multi_device_e2e.jsonl 2> gpurun_out/r1c_multi_device_e2e.err; echo "mde exit $?")
cat gpurun_out/r1c_multi_device_e2e.jsonl; tail -3 gpurun_out/r1c_multi_device_e2e.err
(timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r1c_bench_n2.json 2> gpurun_out/r1c_bench_n2.err; echo "bench2 exit $?")
tail -c 600 gpurun_out/r1c_bench_n2.json
