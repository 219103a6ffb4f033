set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r2_gpu.txt
timeout 1800 python -m pytest tests -m gpu -x -q --timeout 900 > gpurun_out/r2_pytest1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest1.log
tail -5 gpurun_out/r2_pytest1.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench1.log 2> gpurun_out/r2_bench1.err; echo "bench rc=$?"
tail -c 400 gpurun_out/r2_bench1.err
# launch list of one step (shares)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 18 -c 28 --csv --log-file gpurun_out/r2_launches1.csv python tools/profile_step.py c2_1M_1080p_sh3 3 > gpurun_out/r2_launches1.out 2>&1
# full capture of the binning + blend kernels, C2 and C5 (second step of two)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"rasterize_|sort_pack|bin_count|count_scan|bucket_emit|tile_scan|reduce_grad" -s 13 -c 9 -o gpurun_out/r2_prof_c2 python tools/profile_step.py c2_1M_1080p_sh3 2 > gpurun_out/r2_prof_c2.out 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"rasterize_|sort_pack" -s 5 -c 4 -o gpurun_out/r2_prof_c5 python tools/profile_step.py c5_5M_1440p_dense 2 > gpurun_out/r2_prof_c5.out 2>&1
ls -la gpurun_out | tail -12
# A/B: per-tile sort: distribution sort (default) vs bitonic only vs CTA radix above 512
for lib in default opensplat_b200/lib/variants/lib_nodsort.so opensplat_b200/lib/variants/lib_nodsort_radix512.so; do
  for wl in c2_1M_1080p_sh3 c5_5M_1440p_dense; do
    if [ "$lib" = default ]; then timeout 300 python tools/bench_blend.py $wl 10 >> gpurun_out/r2_ab_sort.log 2>&1
    else GSB_LIB=$lib timeout 300 python tools/bench_blend.py $wl 10 >> gpurun_out/r2_ab_sort.log 2>&1; fi
  done
done
cat gpurun_out/r2_ab_sort.log
timeout 900 python tools/bench_model_train.py --steps 20 > gpurun_out/r2_model_train.json 2> gpurun_out/r2_model_train.err; tail -c 600 gpurun_out/r2_model_train.json
timeout 120 opensplat_b200/lib/ubench_fp32 > gpurun_out/r2_ubench_fp32.txt 2>&1; cat gpurun_out/r2_ubench_fp32.txt
# sanitizers on the new kernels (binning v2, distribution sort, exchange launch): smoke() + the exchange tests
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_san_memcheck_smoke.txt 2>&1; tail -3 gpurun_out/r2_san_memcheck_smoke.txt
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_exchange.py -x -q > gpurun_out/r2_san_memcheck_exchange.txt 2>&1; tail -3 gpurun_out/r2_san_memcheck_exchange.txt
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_san_racecheck_smoke.txt 2>&1; tail -3 gpurun_out/r2_san_racecheck_smoke.txt
