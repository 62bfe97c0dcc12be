#!/bin/bash
# GPU box session A: ncu --set full of the benchmarked kernel (raw csv kept), then the whole GPU test suite.
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none -k regex:zodi_los -s 2 -c 1"
(timeout 150 $NCU -o gpurun_out/r1b_x2_planck18_nside2048 python benchmarks/profile_target.py --name planck18 --x 857 --unit GHz --nside 2048 > gpurun_out/ncu_x2b.log 2>&1; tail -n 1 gpurun_out/ncu_x2b.log)
ncu -i gpurun_out/r1b_x2_planck18_nside2048.ncu-rep --page raw --csv > gpurun_out/r1b_x2_planck18_nside2048.raw.csv 2>/dev/null
rm -f gpurun_out/r1b_x2_planck18_nside2048.ncu-rep
(timeout 150 $NCU -o gpurun_out/r1b_x2_dirbe_nside1024 python benchmarks/profile_target.py --name dirbe --x 25 --unit um --nside 1024 > gpurun_out/ncu_x2c.log 2>&1; tail -n 1 gpurun_out/ncu_x2c.log)
ncu -i gpurun_out/r1b_x2_dirbe_nside1024.ncu-rep --page raw --csv > gpurun_out/r1b_x2_dirbe_nside1024.raw.csv 2>/dev/null
rm -f gpurun_out/r1b_x2_dirbe_nside1024.ncu-rep
(timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r1b_gpu_tests.log 2>&1; echo "exit $?" >> gpurun_out/r1b_gpu_tests.log)
tail -4 gpurun_out/r1b_gpu_tests.log
