set -x
mkdir -p gpurun_out
N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/r2_topo8.txt 2>&1
timeout 400 $TR --master-port 29711 tools/check_multigpu.py > gpurun_out/r2_mgcheck_n8.log 2>&1; echo "rc=$?" >> gpurun_out/r2_mgcheck_n8.log
grep -v "^\[W\|^$" gpurun_out/r2_mgcheck_n8.log | tail -5
timeout 600 $TR --master-port 29714 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_n8_fused.log 2> gpurun_out/r2_bench_n8_fused.err; echo "rc=$?"
tail -c 300 gpurun_out/r2_bench_n8_fused.err
timeout 400 $TR --master-port 29713 tools/bench_exchange.py > gpurun_out/r2_exchange_n8.json 2> gpurun_out/r2_exchange_n8.err; tail -c 1800 gpurun_out/r2_exchange_n8.json
timeout 600 $TR --master-port 29715 bench.py --gpus $N --steps 20 --warmup 5 --exchange nccl > gpurun_out/r2_bench_n8_nccl.log 2> gpurun_out/r2_bench_n8_nccl.err; echo "rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench_n8_*.log')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['stages_ms'], d.get('exchange_check'), [ (p['rank'],p['intersections_binned'],p['compute_ms_without_exchange'],p['exchange_stage_ms']) for p in d['per_rank']])
    except Exception as e: print(f, 'ERR', e)
PY
