export WG_B200_DEV_LIB=1   # the knobs below exist in the developer build only (make dev)
tools/selftest tsrate > gpurun_out/tsrate.txt 2>&1
for d in 0 1 2 3 4 8 12 16 6 14 15 31; do
  echo "== WG_FF_DEBUG=$d" 
  WG_FF_DEBUG=$d timeout 120 python tools/ff_check.py --time-only --kns 96 --dtypes tf32 --iters 30 --out gpurun_out/abl_$d.json 2>&1 | grep "'time'" | sed -e "s/'check': 'time', //" -e "s/'rel_err.*//"
done
