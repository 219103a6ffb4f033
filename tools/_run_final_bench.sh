set -x
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2_bench_final.log 2> gpurun_out/r2_bench_final.err; echo "rc=$?"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python - <<'PY'
import json
s=open('gpurun_out/r2_bench_final.log').read(); d=json.loads(s[s.index('{'):])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('train_iters_per_s'), d['stages_ms'], d['clocks'], d['gpu_launches'])
print({k:(v.get('value'),v.get('ms_per_step')) for k,v in d.get('other_configs',{}).items()})
print(d['roofline']['frac'], d['roofline']['traffic'], d['cpu_baseline'])
PY
