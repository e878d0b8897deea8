export WG_B200_DEV_LIB=1
for a in ${ONE_TL_MODES:-16}; do for s in "256 1024" "128 512" "512 128" "1024 256"; do WG_ONE_ABLATE=$a python tools/one_timeline.py $s; done; done
