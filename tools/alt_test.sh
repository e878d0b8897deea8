export WG_B200_DEV_LIB=1
for a in 0 1; do for d in 128 132 136 140 129 130; do
  echo "##### WG_FF_ALT=$a WG_FF_DEBUG=$d"
  WG_FF_ALT=$a WG_FF_DEBUG=$d python tools/quick.py --ns 256 --iters 30 --tag "alt=$a dbg=$d" | grep "128->128"
done; done
