set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_ff.log 2>&1; echo "pytest rc=$?" 
tail -5 gpurun_out/pytest_gpu_ff.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_ff.json 2> gpurun_out/bench_ff.err; echo "bench rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_ff.csv python bench.py --steps 20 --warmup 3 > gpurun_out/ncu_launches.log 2>&1; echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:wino3x3_ff_kernel --launch-skip 5 --launch-count 1 -o gpurun_out/prof_ff_r01 python bench.py --steps 8 --warmup 3 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/*.ncu-rep
