# ncu --set full capture of the 3x3 throughput kernel (one launch, after warm-up) for C=K=256 and C=K=128 at N=256,
# rotating buffers; the same command exits 0 without ncu first.
set -x
mkdir -p gpurun_out
python tools/quick.py --ns 256 --iters 10 --tag pre-ncu || exit 1
ncu --set full --clock-control none --import-source on -k regex:wino3x3_ff_kernel --launch-skip 6 --launch-count 1 -f -o gpurun_out/prof_ff256_r02 python tools/quick.py --ns 256 --iters 4 > gpurun_out/ncu_ff256.log 2>&1; echo "ncu rc=$?"
ncu --set full --clock-control none --import-source on -k regex:wino3x3_ff_kernel --launch-skip 40 --launch-count 1 -f -o gpurun_out/prof_ff128_r02 python tools/quick.py --ns 256 --iters 4 > gpurun_out/ncu_ff128.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep
