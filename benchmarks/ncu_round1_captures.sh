# ncu --set full captures of the round-1 kernels (run under gpurun, 1 GPU).  Raw metric pages are
# exported on the box (gpurun_out/ is limited to 64 MiB); only two reports travel back.
set -x
NCU="ncu --set full --clock-control none -k regex:zodi_los -s 2 -c 1"
T="python benchmarks/profile_target.py"
$NCU -o gpurun_out/r1_x2_planck18_nside2048 $T --name planck18 --x 857 --unit GHz --nside 2048 > gpurun_out/ncu_a.log 2>&1
$NCU --import-source on -o gpurun_out/r1_fp64_planck18_nside1024 $T --name planck18 --x 857 --unit GHz --nside 1024 --precision fp64 > gpurun_out/ncu_b.log 2>&1
$NCU --import-source on -o gpurun_out/r1_x2_scatter_dirbe1p25_nside1024 $T --x 1.25 --unit um --nside 1024 > gpurun_out/ncu_c.log 2>&1
$NCU -o gpurun_out/r1_generic_rrm_nside512 $T --name rrm-experimental --x 25 --unit um --nside 512 > gpurun_out/ncu_d.log 2>&1
for r in gpurun_out/r1_x2_planck18_nside2048 gpurun_out/r1_fp64_planck18_nside1024 gpurun_out/r1_x2_scatter_dirbe1p25_nside1024 gpurun_out/r1_generic_rrm_nside512; do
  ncu -i $r.ncu-rep --page raw --csv > $r.raw.csv
done
rm -f gpurun_out/r1_x2_planck18_nside2048.ncu-rep gpurun_out/r1_generic_rrm_nside512.ncu-rep
for f in gpurun_out/ncu_a.log gpurun_out/ncu_b.log gpurun_out/ncu_c.log gpurun_out/ncu_d.log; do tail -n 2 $f; done
du -sh gpurun_out
