#!/bin/bash
# Round-2 GPU pass G (1 GPU): CTA-size sweep at shard sizes (tail of a launch), parity, launch list + bench line.
mkdir -p gpurun_out
timeout 600 python benchmarks/n_sweep.py --shapes --models planck18,dirbe --nsides 256,512,724,1024 > gpurun_out/r2g_sweep.jsonl 2> gpurun_out/r2g_sweep.err
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
tail -3 gpurun_out/r2g_pytest.log
python - <<'PY'
import json
for l in open('gpurun_out/r2g_sweep.jsonl'):
    d=json.loads(l); print(d['model'],d['nside'],d['lanes'],d['threads'],'%.4f'%d['ms'],'%.3e'%d['evals_per_s'])
PY
