# N-GPU round (N = number of visible GPUs): bench line through torchrun, C example
set -x
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_r02_${N}gpu.json 2> gpurun_out/bench_r02_${N}gpu.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_r02_${N}gpu.err
./examples/shard_batch 256 256 | tee gpurun_out/shard_batch_${N}gpu.txt
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r02_${N}gpu.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus')})
print('e2e',d['e2e']['value'], d['e2e'].get('pcie_ceiling'))
print(json.dumps(d.get('with_output_allgather'))[:1200])
for r in d['strong_scaling']['rows']: print(r)
for r in d['bottleneck_block']: print(r)
PY
