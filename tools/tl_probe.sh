export WG_B200_DEV_LIB=1
for d in ${TL_MODES:-112 113}; do for c in ${TL_CS:-256 128}; do echo "#### WG_FF_DEBUG=$d C=$c"; WG_FF_DEBUG=$d python tools/ff_timeline.py $c tf32only 2>&1 | grep -v "^$\|sorted (clk:block)\|globaltimer ns\|both items" | head -40; done; done
