export WG_B200_DEV_LIB=1
for a in 31 20 28; do for s in "256 1024" "512 128"; do WG_ONE_ABLATE=$a python tools/one_timeline.py $s; done; done
