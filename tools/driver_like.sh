set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 3 ) 2>&1 | tail -5 | cut -c1-600
( time python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench_default.json ) 2>&1 | tail -4
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_default.json'))
print({k:d[k] for k in ('metric','value','unit','n_gpus','steps','warmup','ms_per_step','higher_is_better','scaling','vs_baseline','dtype','data','gpu_launches')})
print(d['config']); print(d['clocks'])
print({k:v for k,v in d['roofline'].items() if k not in ('traffic_detail',)})
print(d['e2e'])
print(d['cpu_baseline'])
print(list(d.keys()))
PY
