# ncu --set full captures of the round-2 kernels with source-level counters (run under gpurun, 1 GPU).
# Raw metric pages and the per-SASS-instruction source page are exported on the box (gpurun_out/ is
# limited to 64 MiB); the reports themselves stay there.
# Usage: bash benchmarks/ncu_round2_captures.sh [x2|dirbe|fp64|rrm|dirbe64 ...]   (default: x2 dirbe)
NCU="ncu --set full --clock-control none --import-source on -k regex:zodi_los -s 2 -c 1"
T="python benchmarks/profile_target.py"
WHAT="${@:-x2 dirbe}"
TAG="${NCU_TAG:-r2}"
for w in $WHAT; do
  case $w in
    x2) name=${TAG}_x2_planck18_nside2048; args="--name planck18 --x 857 --unit GHz --nside 2048";;
    x2a) name=${TAG}_x2_planck18_nside2048_arrays; args="--name planck18 --x 857 --unit GHz --nside 2048 --arrays";;
    fp64a) name=${TAG}_fp64_planck18_nside2048_arrays; args="--name planck18 --x 857 --unit GHz --nside 2048 --precision fp64 --arrays";;
    dirbe) name=${TAG}_x2_dirbe_nside1024; args="--name dirbe --x 25 --unit um --nside 1024";;
    dirbe64) name=${TAG}_x2_dirbe_nside64; args="--name dirbe --x 25 --unit um --nside 64";;
    fp64) name=${TAG}_fp64_planck18_nside1024; args="--name planck18 --x 857 --unit GHz --nside 1024 --precision fp64";;
    rrm) name=${TAG}_rrm_nside512; args="--name rrm-experimental --x 25 --unit um --nside 512";;
    mb) name=${TAG}_multiband_x2_dirbe10_nside512; args="--name dirbe --unit um --nside 512 --bands 1.25,2.2,3.5,4.9,12,25,60,100,140,240";;
    mbp) name=${TAG}_multiband_x2_planck6_nside1024; args="--name planck18 --unit GHz --nside 1024 --bands 100,143,217,353,545,857";;
    rrm64) name=${TAG}_rrm_fp64_nside256; args="--name rrm-experimental --x 25 --unit um --nside 256 --precision fp64";;
  esac
  $NCU -o gpurun_out/$name $T $args > gpurun_out/ncu_$w.log 2>&1
  ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv
  ncu -i gpurun_out/$name.ncu-rep --page source --csv --print-source sass > gpurun_out/$name.source.csv 2> gpurun_out/ncu_src_$w.err
  rm -f gpurun_out/$name.ncu-rep
  tail -n 1 gpurun_out/ncu_$w.log
done
