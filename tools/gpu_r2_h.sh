#!/bin/bash
# Round-2 GPU pass H (1 GPU): full parity suite.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/r2h_pytest.log | tail -15
