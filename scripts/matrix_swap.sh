export TRB_TC_ISSUERS_HALO=1
for h in 3 1 2; do echo "== tests TRB_TC_ISSUERS=2 TRB_TC_HALO=$h: $(TRB_TC_ISSUERS=2 TRB_TC_HALO=$h timeout 120 python -m pytest tests/test_gpu_ops.py -q -m gpu --tb=line -k 'tcgen05 or fp32' 2>&1 | tail -1)"; done
echo "== microbench halo issuers ON"; timeout 120 python scripts/bench_conv.py "vgg 3x3 64" "vgg 3x3 128->128" "arcface 3x3 64" "arcface 3x3 128" "retina 3x3"
echo "== microbench halo issuers OFF"; TRB_TC_ISSUERS_HALO=0 timeout 120 python scripts/bench_conv.py "vgg 3x3 64" "vgg 3x3 128->128" "arcface 3x3 64" "arcface 3x3 128" "retina 3x3"
echo "== nets halo issuers ON"; timeout 200 python scripts/profile_ops.py openpose arcface --brief | grep -E "^==|rror"
