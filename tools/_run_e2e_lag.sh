set -x
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r2_bench_lag_n1.log 2> gpurun_out/r2_bench_lag_n1.err; echo "rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29817 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_lag_n2.log 2> gpurun_out/r2_bench_lag_n2.err; echo "rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench_lag_n1.log", "gpurun_out/r2_bench_lag_n2.log"):
    s=open(f).read(); d=json.loads(s[s.index('{'):])
    print(f, d['n_gpus'], d['value'], d['ms_per_step'], d['e2e'], d.get('exchange_check'))
PY
