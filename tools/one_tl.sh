export WG_B200_DEV_LIB=1
for m in ${ONE_TL_MODES:-16}; do for s in "128 512" "256 1024"; do WG_ONE_ABLATE=$m python tools/one_timeline.py $s 256 > gpurun_out/otl.txt 2>&1; grep -A10 "epilogue warp 2\|==" gpurun_out/otl.txt | grep -v "CTAs with\|producer" | head -24; done; done
