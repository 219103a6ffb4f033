set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2_pytest7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest7.log
tail -3 gpurun_out/r2_pytest7.log
for rep in 1 2; do
  timeout 200 python tools/bench_blend.py c2_1M_1080p_sh3 10 >> gpurun_out/r2_ab_occupancy.log 2>&1
  for v in bwd7 bwd5 fwd7 fwd6; do
    GSB_LIB=opensplat_b200/lib/variants/lib_$v.so timeout 200 python tools/bench_blend.py c2_1M_1080p_sh3 10 >> gpurun_out/r2_ab_occupancy.log 2>&1
  done
done
for v in default bwd7 fwd7; do
  if [ $v = default ]; then timeout 200 python tools/bench_blend.py c5_5M_1440p_dense 6 >> gpurun_out/r2_ab_occupancy.log 2>&1
  else GSB_LIB=opensplat_b200/lib/variants/lib_$v.so timeout 200 python tools/bench_blend.py c5_5M_1440p_dense 6 >> gpurun_out/r2_ab_occupancy.log 2>&1; fi
done
sed 's/sh_fwd=.*raster_fwd=/raster_fwd=/; s/loss=[0-9.]* //; s/project_bwd.*//' gpurun_out/r2_ab_occupancy.log
