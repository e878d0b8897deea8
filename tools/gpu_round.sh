# One GPU round for the profiles/ directory: tests, bench line, ncu launch list + full capture of the headline kernel,
# chain benchmark, probes. Everything lands in gpurun_out/.
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r01_ff.json 2> gpurun_out/bench_ff.err; echo "bench rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_ff.csv python bench.py --steps 20 --warmup 3 > gpurun_out/ncu_launches.log 2>&1; echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:wino3x3_ff_kernel --launch-skip 5 --launch-count 1 -o gpurun_out/prof_ff_r01 python bench.py --steps 8 --warmup 3 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
python tools/chain_bench.py > gpurun_out/chain_bench_r01_ff.json 2> gpurun_out/chain.err; echo "chain rc=$?"
python tools/ff_timeline.py > gpurun_out/ff_timeline_r01.txt 2>&1
tools/selftest lds > gpurun_out/lds_probe_r01.txt 2>&1
python tools/ff_check.py --time-only --iters 40 --out gpurun_out/ff_check_times_r01.json > gpurun_out/ff_check_times.log 2>&1
