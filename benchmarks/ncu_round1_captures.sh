# ncu --set full captures of the round-1 kernels (run under gpurun, 1 GPU).  Raw metric pages are
# exported on the box (gpurun_out/ is limited to 64 MiB); reports themselves stay there unless listed
# in KEEP.  Usage: bash benchmarks/ncu_round1_captures.sh [fp64|x2|scatter|rrm ...]   (default: all)
set -x
NCU="ncu --set full --clock-control none -k regex:zodi_los -s 2 -c 1"
T="python benchmarks/profile_target.py"
WHAT="${@:-x2 fp64 scatter rrm}"
for w in $WHAT; do
  case $w in
    x2) name=r1_x2_planck18_nside2048; args="--name planck18 --x 857 --unit GHz --nside 2048";;
    fp64) name=r1_fp64_planck18_nside1024; args="--name planck18 --x 857 --unit GHz --nside 1024 --precision fp64";;
    scatter) name=r1_x2_scatter_dirbe1p25_nside1024; args="--x 1.25 --unit um --nside 1024";;
    rrm) name=r1_generic_rrm_nside512; args="--name rrm-experimental --x 25 --unit um --nside 512";;
  esac
  $NCU -o gpurun_out/$name $T $args > gpurun_out/ncu_$w.log 2>&1
  ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv
  rm -f gpurun_out/$name.ncu-rep
  tail -n 1 gpurun_out/ncu_$w.log
done
