set -x
mkdir -p gpurun_out
N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29711 tools/check_multigpu.py > gpurun_out/r2_mgcheck_n8b.log 2>&1; echo "rc=$?" >> gpurun_out/r2_mgcheck_n8b.log
grep -v "^\[W\|^$" gpurun_out/r2_mgcheck_n8b.log | tail -5
timeout 600 $TR --master-port 29714 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_n8_overlap.log 2> gpurun_out/r2_bench_n8_overlap.err; echo "rc=$?"
tail -c 300 gpurun_out/r2_bench_n8_overlap.err
GSB_EXCHANGE_OVERLAP=0 timeout 600 $TR --master-port 29715 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_n8_single.log 2> gpurun_out/r2_bench_n8_single.err; echo "rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29716 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2_bench_n4_overlap.log 2> gpurun_out/r2_bench_n4_overlap.err; echo "rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench_n[48]_*.log')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['stages_ms'], d.get('exchange_check'), [ (p['rank'],p['intersections_binned'],p['compute_ms_without_exchange'],p['exchange_stage_ms']) for p in d['per_rank']])
    except Exception as e: print(f, 'ERR', e)
PY
