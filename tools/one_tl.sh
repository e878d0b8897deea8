export WG_B200_DEV_LIB=1
for s in "1024 256" "512 128"; do WG_ONE_ABLATE=${ONE_TL_MODE:-16} python tools/one_timeline.py $s 256 > gpurun_out/otl.txt 2>&1; head -8 gpurun_out/otl.txt; done
