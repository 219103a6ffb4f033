set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2_pytest5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest5.log
tail -4 gpurun_out/r2_pytest5.log
for rep in 1 2; do
for wl in c2_1M_1080p_sh3 c5_5M_1440p_dense c3_3M_4k_sh3; do
  timeout 300 python tools/bench_blend.py $wl 10 >> gpurun_out/r2_ab_gpos.log 2>&1
  GSB_LIB=opensplat_b200/lib/variants/lib_prev.so timeout 300 python tools/bench_blend.py $wl 10 >> gpurun_out/r2_ab_gpos.log 2>&1
done; done
cat gpurun_out/r2_ab_gpos.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench5.log 2> gpurun_out/r2_bench5.err; echo "bench rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sort_pack|bucket_emit" -s 3 -c 3 -o gpurun_out/r2_prof_c5b python tools/profile_step.py c5_5M_1440p_dense 2 > gpurun_out/r2_prof_c5b.out 2>&1
