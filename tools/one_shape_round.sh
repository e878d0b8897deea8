for s in "256 1024 --relu 0" "1024 256" "128 512 --relu 0" "512 128"; do python tools/one_shape.py 1x1 $s; done
ncu --set full --clock-control none --import-source on -k regex:conv1x1_bn_act --launch-skip 3 --launch-count 1 -o gpurun_out/prof_1x1_256_1024 python tools/one_shape.py 1x1 256 1024 --relu 0 --iters 5 > gpurun_out/ncu_1x1.log 2>&1; echo ncu rc=$?
