#!/bin/bash
# Round-2 GPU pass R (1 GPU): packed multi-band kernel with host-transposed table rows - parity and timing.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multiband" > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2r_pytest.log
timeout 200 python benchmarks/baseline_configs.py --multiband-only > gpurun_out/r2r_multiband.jsonl 2> gpurun_out/r2r_multiband.err
tail -3 gpurun_out/r2r_pytest.log
python - <<'PY'
import json
for l in open('gpurun_out/r2r_multiband.jsonl'):
    d=json.loads(l)
    print('  ', d['config'], 'packed %.4f scalar %.4f loop %.4f'%(d['multiband_ms'], d['scalar_multiband_ms'], d['per_band_loop_ms']), 'diff %.2e'%d['max_rel_diff_vs_single_band'])
PY
tail -c 300 gpurun_out/r2r_multiband.err
