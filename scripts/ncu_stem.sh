python scripts/profile_ops.py retinaface openpose arcface --brief | grep -E "^==|stem"
ncu --set full --clock-control none --import-source on -k regex:stem_mma -s 3 -c 1 -f -o gpurun_out/prof_stem_1 python scripts/profile_ops.py openpose --brief > gpurun_out/ncu_stem_1.log 2>&1
