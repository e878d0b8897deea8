export WG_B200_DEV_LIB=1
for s in "1024 256" "512 128" "256 1024" "128 512"; do WG_ONE_ABLATE=16 python tools/one_timeline.py $s 256 | head -6; done
