set -x
mkdir -p gpurun_out
N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
P=29720
for gb in 4 1 2 8 16; do
  P=$((P+1))
  GSB_GEOM_BLOCKS=$gb timeout 300 $TR --master-port $P tools/bench_exchange.py --reps 20 > gpurun_out/r2_exchange_n8_gb$gb.json 2> gpurun_out/r2_exchange_n8_gb$gb.err
  python - <<PY
import json
s=open('gpurun_out/r2_exchange_n8_gb$gb.json').read(); d=json.loads(s[s.index('{'):])
for f in ('multimem','peer'):
    r=d[f]; print('gb=$gb',f,{k:round(v,4) for k,v in r.items() if k.endswith('_ms')})
print('nccl',d['nccl_flat_allreduce'])
PY
done
