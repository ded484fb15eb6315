for c in 1024 512; do echo "== TRB_TC_STAGE_CYCLES=$c"; TRB_TC_STAGE_CYCLES=$c python scripts/bench_conv.py; done
for c in 1024 512 1024 512; do echo "== net TRB_TC_STAGE_CYCLES=$c"; TRB_TC_STAGE_CYCLES=$c python scripts/profile_ops.py openpose arcface retinaface --brief | grep -E "^==|tcgen05"; done
