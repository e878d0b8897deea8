export WG_B200_DEV_LIB=1   # the knobs below exist in the developer build only (make dev)
# A/B runs of the full-fold kernel's knobs (read once per process): prints the N=256 times
run() { echo "== $*"; env "$@" timeout 300 python tools/ff_check.py --quick --kns 96 --iters 40 --out gpurun_out/ab.json 2>&1 | grep "'time'\|failures" | sed -e "s/'check': 'time', 'kn': 96, //" -e "s/, 'rel_err.*//"; }
run WG_FF_CG2=1 WG_FF_W16=1
run WG_FF_CG2=1 WG_FF_W16=0
run WG_FF_CG2=0 WG_FF_W16=1
run WG_FF_CG2=0 WG_FF_W16=0
