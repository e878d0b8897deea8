# ncu --set full captures of the 1x1 throughput kernel, all four README shapes at N=256 (quick.py runs them in the order
# 512->128, 128->512, 1024->256, 256->1024; each shape: 4 warm-up + 2 x iters plain launches, then the residual ones)
set -x
mkdir -p gpurun_out
python tools/quick.py --ns 256 --one --iters 4 --tag pre-ncu > /dev/null || exit 1
ncu --set full --clock-control none --import-source on -k regex:conv1x1_bn_act_kernel --launch-skip 6 --launch-count 1 -f -o gpurun_out/prof_one_512_128_r02 python tools/quick.py --ns 256 --one --iters 4 > gpurun_out/ncu_one_a.log 2>&1; echo "ncu rc=$?"
ncu --set full --clock-control none --import-source on -k regex:conv1x1_bn_act_kernel --launch-skip 30 --launch-count 1 -f -o gpurun_out/prof_one_128_512_r02 python tools/quick.py --ns 256 --one --iters 4 > gpurun_out/ncu_one_b.log 2>&1; echo "ncu rc=$?"
ncu --set full --clock-control none --import-source on -k regex:conv1x1_bn_act_kernel --launch-skip 54 --launch-count 1 -f -o gpurun_out/prof_one_1024_256_r02 python tools/quick.py --ns 256 --one --iters 4 > gpurun_out/ncu_one_c.log 2>&1; echo "ncu rc=$?"
ncu --set full --clock-control none --import-source on -k regex:conv1x1_bn_act_kernel --launch-skip 78 --launch-count 1 -f -o gpurun_out/prof_one_256_1024_r02 python tools/quick.py --ns 256 --one --iters 4 > gpurun_out/ncu_one_d.log 2>&1; echo "ncu rc=$?"
ncu --set full --clock-control none --import-source on -k regex:wino3x3_ff_kernel --launch-skip 18 --launch-count 1 -f -o gpurun_out/prof_ff128_r02 python tools/quick.py --ns 256 --iters 4 > gpurun_out/ncu_ff128.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep
