# One GPU round for profiles/ (round 2): tests, bench line, ncu launch list of the bench command, ncu --set full captures of
# the headline kernel, the 128->128 3x3 kernel and the four 1x1 shapes. Everything lands in gpurun_out/. Each ncu command
# runs only after the same command has exited 0 without ncu.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu_r02.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_r02.log
python tools/sanitize_cases.py > gpurun_out/sanitize_plain.log 2>&1; echo "sanitize_cases rc=$?"
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err; echo "bench rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 20 --warmup 3 > gpurun_out/ncu_launches.log 2>&1; echo "ncu list rc=$?"
python tools/quick.py --ns 256 --one --iters 4 --tag pre-ncu > /dev/null || exit 1
Q="python tools/quick.py --ns 256 --one --iters 4"
NCU="ncu --set full --clock-control none --import-source on --launch-count 1 -f"
$NCU -k regex:conv3x3_direct_kernel --launch-skip 6 -o gpurun_out/prof_dir256_r02 $Q > gpurun_out/ncu_a.log 2>&1; echo "ncu rc=$?"
$NCU -k regex:conv3x3_direct_kernel --launch-skip 18 -o gpurun_out/prof_dir128_r02 $Q > gpurun_out/ncu_b.log 2>&1; echo "ncu rc=$?"
$NCU -k regex:conv1x1_bn_act_kernel --launch-skip 6 -o gpurun_out/prof_one_512_128_r02 $Q > gpurun_out/ncu_c.log 2>&1; echo "ncu rc=$?"
$NCU -k regex:conv1x1_t_kernel --launch-skip 6 -o gpurun_out/prof_one_128_512_r02 $Q > gpurun_out/ncu_d.log 2>&1; echo "ncu rc=$?"
$NCU -k regex:conv1x1_bn_act_kernel --launch-skip 30 -o gpurun_out/prof_one_1024_256_r02 $Q > gpurun_out/ncu_e.log 2>&1; echo "ncu rc=$?"
$NCU -k regex:conv1x1_t_kernel --launch-skip 30 -o gpurun_out/prof_one_256_1024_r02 $Q > gpurun_out/ncu_f.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep
