export WG_B200_DEV_LIB=1
for a in 0 1 2 3 4 8 5 7 15; do WG_ONE_ABLATE=$a python tools/quick.py --ns 256 --one --iters 40 --tag "ablate=$a" | grep -v residual | grep "1x1\|quick"; done
