set -x
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29611 tools/check_multigpu.py > gpurun_out/r2_mgcheck_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/r2_mgcheck_n$N.log
grep -v "^\[W\|^$" gpurun_out/r2_mgcheck_n$N.log | tail -6
GSB_EXCHANGE_MULTICAST=0 timeout 600 $TR --master-port 29612 tools/check_multigpu.py > gpurun_out/r2_mgcheck_peer_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/r2_mgcheck_peer_n$N.log
grep -v "^\[W\|^$" gpurun_out/r2_mgcheck_peer_n$N.log | tail -4
timeout 600 $TR --master-port 29613 tools/bench_exchange.py > gpurun_out/r2_exchange_n$N.json 2> gpurun_out/r2_exchange_n$N.err; tail -c 1500 gpurun_out/r2_exchange_n$N.json
timeout 900 $TR --master-port 29614 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_n${N}_fused.log 2> gpurun_out/r2_bench_n${N}_fused.err; echo "rc=$?"
tail -c 300 gpurun_out/r2_bench_n${N}_fused.err
timeout 900 $TR --master-port 29615 bench.py --gpus $N --steps 20 --warmup 5 --exchange nccl > gpurun_out/r2_bench_n${N}_nccl.log 2> gpurun_out/r2_bench_n${N}_nccl.err; echo "rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench_n*_*.log')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['stages_ms'], d.get('exchange_check'), [ (p['rank'],p['intersections_binned'],p['compute_ms_without_exchange'],p['exchange_stage_ms']) for p in d['per_rank']])
    except Exception as e: print(f, 'ERR', e)
PY
