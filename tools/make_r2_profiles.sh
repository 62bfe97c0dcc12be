#!/bin/bash
# Regenerates profiles/r2_ncu_*.md and profiles/kernel_counts.json (read by bench.py) from the raw metric
# pages exported by benchmarks/ncu_round2_captures.sh (NCU_TAG=r2k) into gpurun_out/.
set -e
S="python tools/ncu_summary.py"
C="--counts-json profiles/kernel_counts.json"
CMD="ncu --set full --clock-control none --import-source on -k regex:zodi_los -s 2 -c 1 python benchmarks/profile_target.py"
$S gpurun_out/r2k_x2_planck18_nside2048_arrays.raw.csv \
  --title "ncu --set full, zodi_los_kelsall_x2_kernel<cloud+bands, L=1, 128-thread CTAs>, planck18 857 GHz nside 2048 fp32 from device arrays (bench.py's value path), round-2 final build" \
  --command "$CMD --name planck18 --x 857 --unit GHz --nside 2048 --arrays" \
  --workload "50 331 648 lines of sight x 4 comps x 50 nodes = 1.0066e10 evaluations in one launch, (3, N) float64 unit vectors read from HBM" \
  --los 50331648 --evals 1.00663296e10 $C --counts-key planck18_fp32_packed --source profiles/r2_ncu_x2_planck18_nside2048_arrays.md \
  > profiles/r2_ncu_x2_planck18_nside2048_arrays.md
$S gpurun_out/r2k_fp64_planck18_nside2048_arrays.raw.csv \
  --title "ncu --set full, zodi_los_kelsall_kernel<double, cloud+bands, L=1>, planck18 857 GHz nside 2048 fp64 (faithful mode) from device arrays, round-2 final build" \
  --command "$CMD --name planck18 --x 857 --unit GHz --nside 2048 --precision fp64 --arrays" \
  --workload "50 331 648 lines of sight x 4 comps x 50 nodes = 1.0066e10 evaluations in one launch" \
  --los 50331648 --evals 1.00663296e10 $C --counts-key planck18_fp64 --source profiles/r2_ncu_fp64_planck18_nside2048_arrays.md \
  > profiles/r2_ncu_fp64_planck18_nside2048_arrays.md
$S gpurun_out/r2k_x2_dirbe_nside1024.raw.csv \
  --title "ncu --set full, zodi_los_kelsall_x2_kernel<cloud+bands, ring|ring, feature|feature, L=1, 128-thread CTAs>, dirbe 25 um nside 1024 fp32, round-2 final build" \
  --command "$CMD --name dirbe --x 25 --unit um --nside 1024" \
  --workload "12 582 912 lines of sight x 6 comps x 50 nodes = 3.7749e9 evaluations in one launch, pixel directions generated in the kernel prologue" \
  --los 12582912 --evals 3.7748736e9 $C --counts-key dirbe_fp32_packed --source profiles/r2_ncu_x2_dirbe_nside1024.md \
  > profiles/r2_ncu_x2_dirbe_nside1024.md
RRM=${RRM_RAW:-gpurun_out/r2k_rrm_nside512.raw.csv}
$S $RRM \
  --title "ncu --set full, fused RRM kernel, rrm-experimental 25 um nside 512 fp32, round 2" \
  --command "$CMD --name rrm-experimental --x 25 --unit um --nside 512" \
  --workload "3 145 728 lines of sight x 8 comps x 50 nodes = 1.2583e9 evaluations in one launch" \
  --los 3145728 --evals 1.2582912e9 $C --counts-key rrm_fp32 --source profiles/r2_ncu_rrm_nside512.md \
  > profiles/r2_ncu_rrm_nside512.md
for f in x2_planck18_nside2048_arrays fp64_planck18_nside2048_arrays x2_dirbe_nside1024; do
  python tools/ncu_source_hot.py gpurun_out/r2k_$f.source.csv --top 12 > profiles/r2_ncu_${f}_source_hot.txt
done
# packed multi-band kernel (pass Q capture)
if [ -f gpurun_out/r2q_multiband_x2_planck6_nside1024.raw.csv ]; then
  $S gpurun_out/r2q_multiband_x2_planck6_nside1024.raw.csv \
    --title "ncu --set full, zodi_los_multiband_x2_kernel<NB=8, cloud+bands>, planck18 6 channels nside 1024 fp32, round-2 final build" \
    --command "$CMD --name planck18 --unit GHz --nside 1024 --bands 100,143,217,353,545,857" \
    --workload "12 582 912 lines of sight x 6 bands x 4 comps x 50 nodes = 1.5099e10 band-evaluations in one launch, pixel directions generated in the kernel prologue" \
    --los 12582912 --evals 1.50994944e10 $C --counts-key planck18_multiband6_fp32_packed --source profiles/r2_ncu_multiband_x2_planck6_nside1024.md \
    > profiles/r2_ncu_multiband_x2_planck6_nside1024.md
  python tools/ncu_source_hot.py gpurun_out/r2q_multiband_x2_planck6_nside1024.source.csv --top 12 > profiles/r2_ncu_multiband_x2_planck6_nside1024_source_hot.txt
fi
