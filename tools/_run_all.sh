set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r2_gpu.txt
nvidia-smi topo -m > gpurun_out/r2_topo.txt 2>&1
NG=$(nvidia-smi -L | wc -l)
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2_pytest6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest3.log
tail -8 gpurun_out/r2_pytest6.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench6.log 2> gpurun_out/r2_bench6.err; echo "bench rc=$?"
tail -c 400 gpurun_out/r2_bench6.err
if [ "$NG" -ge 2 ]; then bash tools/_run2.sh 2; fi
# A/B: longest-first tile order on/off (stage times of the blend kernels)
for wl in c2_1M_1080p_sh3 c5_5M_1440p_dense; do
  GSB_TILE_ORDER=1 timeout 300 python tools/bench_blend.py $wl 10 >> gpurun_out/r2_ab_tileorder.log 2>&1
  GSB_TILE_ORDER=0 timeout 300 python tools/bench_blend.py $wl 10 >> gpurun_out/r2_ab_tileorder.log 2>&1
done
cat gpurun_out/r2_ab_tileorder.log
# launch list of one step (shares) + full captures (C2: all binning + blend kernels of step 2; C5: blend + sort)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 30 --csv --log-file gpurun_out/r2_launches6.csv python tools/profile_step.py c2_1M_1080p_sh3 3 > gpurun_out/r2_launches6.out 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"rasterize_|sort_pack|bin_count|count_scan|bucket_emit|tile_scan|tile_order|reduce_grad" -s 15 -c 10 -o gpurun_out/r2_prof_c2 python tools/profile_step.py c2_1M_1080p_sh3 2 > gpurun_out/r2_prof_c2.out 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"rasterize_|sort_pack" -s 5 -c 4 -o gpurun_out/r2_prof_c5 python tools/profile_step.py c5_5M_1440p_dense 2 > gpurun_out/r2_prof_c5.out 2>&1
timeout 900 python tools/bench_model_train.py --steps 20 > gpurun_out/r2_model_train.json 2> gpurun_out/r2_model_train.err; tail -c 400 gpurun_out/r2_model_train.json
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_exchange.py -x -q > gpurun_out/r2_san_memcheck_exchange.txt 2>&1; tail -3 gpurun_out/r2_san_memcheck_exchange.txt
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_fuzz.py -x -q -k "0 or 3 or 4 or 9" > gpurun_out/r2_san_memcheck_fuzz.txt 2>&1; tail -3 gpurun_out/r2_san_memcheck_fuzz.txt
ls -la gpurun_out | tail -15
