for r in 1 2 3; do echo "== arcface layers TRB_TC_ISSUERS=1 run $r: $(TRB_TC_ISSUERS=1 python scripts/arcface_layers.py | grep -c 'max err') of 11 ok"; done
echo "iso: $(TRB_TC_ISSUERS=1 python scripts/iso.py | grep -c ok) of 12 ok"
for r in 1 2 3; do echo "== nets TRB_TC_ISSUERS=1 run $r"; TRB_TC_ISSUERS=1 python scripts/profile_ops.py openpose arcface retinaface --brief | grep -E "^==|rror"; done
echo "== nets TRB_TC_ISSUERS=0"; TRB_TC_ISSUERS=0 python scripts/profile_ops.py openpose arcface retinaface --brief | grep -E "^=="
echo "== nets TRB_TC_ISSUERS=1 TRB_TC_SUB=1"; TRB_TC_ISSUERS=1 TRB_TC_SUB=1 python scripts/profile_ops.py openpose --brief | grep -E "^==|rror"
TRB_TC_ISSUERS=1 timeout 900 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -3
TRB_TC_ISSUERS=2 timeout 900 python -m pytest tests/test_gpu_ops.py -q -m gpu --tb=line -k "tcgen05 or fp32 or swap" 2>&1 | tail -3
