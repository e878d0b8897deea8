# GPU call A (round 2): smoke, the NEW tests (residual, bf16 1x1, blob, host pipeline, full-tensor N=256), bench line.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke_r02.log
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x --timeout 600 \
  -k "residual or bf16_operand_variant or blob or run_host or full_batch_256 or restored or rejected or throughput_kernel_variants" \
  > gpurun_out/pytest_new_r02.log 2>&1; echo "pytest new rc=$?"; tail -15 gpurun_out/pytest_new_r02.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r02_a.json 2> gpurun_out/bench_r02_a.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench_r02_a.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02_a.json'))
for k in ('value','ms_per_step','clocks','tensor_peak'): print(k, d.get(k))
print('roofline', {k:v for k,v in d['roofline'].items() if k in ('achieved','peak','frac','frac_of_half_bf16_burst')})
print('e2e', d['e2e'])
for r in d['strong_scaling']['rows']: print(r)
for r in d['bottleneck_block']: print(r)
for r in d.get('all_shapes',[]): print(r)
print(json.dumps(d.get('cudnn_baseline'))[:3000])
print(d.get('cpu_baseline'))
PY
