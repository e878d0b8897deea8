# A/B runs of the full-fold kernel's knobs: prints the N=256 times
run() { echo "== $*"; env "$@" timeout 200 python tools/ff_check.py --time-only --kns 96 --dtypes tf32,bf16 --iters 40 --sets ${SETS:-4} --out gpurun_out/ab.json 2>&1 | grep "'time'" | sed -e "s/'check': 'time', 'kn': 96, //" -e "s/, 'rel_err.*//"; }
run WG_FF_DEBUG=0
run WG_FF_DEBUG=128
run WG_FF_DEBUG=16
run WG_FF_DEBUG=17
run WG_FF_DEBUG=18
