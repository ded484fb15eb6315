for i in 1 0; do echo "== arcface layers TRB_TC_ISSUERS=$i"; TRB_TC_ISSUERS=$i python scripts/arcface_layers.py; done
for i in 1 0 1 0; do echo "== net TRB_TC_ISSUERS=$i"; TRB_TC_ISSUERS=$i python scripts/profile_ops.py openpose arcface retinaface --brief | grep -E "^==|k7|tcgen05"; done
