set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_exchange.py -q --timeout 600 > gpurun_out/r2_pytest_multi.log 2>&1; tail -3 gpurun_out/r2_pytest_multi.log
timeout 900 $TR --master-port 29814 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2_final.log 2> gpurun_out/r2_bench_n2_final.err; echo "rc=$?"
timeout 900 $TR --master-port 29815 bench.py --impl reference --gpus 2 --steps 2 --warmup 0 --ref-budget-s 5 > gpurun_out/r2_bench_n2_ref.log 2> gpurun_out/r2_bench_n2_ref.err; echo "rc=$?"; tail -c 300 gpurun_out/r2_bench_n2_ref.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n2_final.log').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['stages_ms'], d['exchange_check'], d['roofline']['pairs'])
PY
