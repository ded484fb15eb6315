for cfg in "TRB_TC_SK=1" "TRB_TC_SK=0"; do
  echo "== $cfg"
  env $cfg python scripts/profile_ops.py arcface openpose --brief
done
