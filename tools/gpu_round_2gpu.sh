# 2-GPU round: the tests that need two GPUs, the C example, the bench line at N=2
set -x
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests -m gpu -q -x --timeout 900 -k "two_devices or fused_output_allgather or restored or c_example" > gpurun_out/pytest_2gpu_r02.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_2gpu_r02.log
./examples/shard_batch 256 256
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_r02_2gpu.json 2> gpurun_out/bench_r02_2gpu.err; echo "bench rc=$?"; tail -c 400 gpurun_out/bench_r02_2gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02_2gpu.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus')})
print('e2e',d['e2e']['value'], d['e2e'].get('pcie_ceiling'))
print(json.dumps(d.get('with_output_allgather'),indent=0)[:1500])
for r in d['strong_scaling']['rows']: print(r)
for r in d['bottleneck_block']: print(r)
PY
