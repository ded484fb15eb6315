echo "-- split (working tree)"; python scripts/profile_ops.py openpose arcface --brief 2>&1 | grep -E "^==" | cut -c1-170
cp terran_b200/lib/libterran_b200.so /tmp/lib_split.so; cp scripts/debug/lib_base.so.bin terran_b200/lib/libterran_b200.so
echo "-- base (HEAD)"; python scripts/profile_ops.py openpose arcface --brief 2>&1 | grep -E "^==" | cut -c1-170
cp /tmp/lib_split.so terran_b200/lib/libterran_b200.so
echo "-- split again"; python scripts/profile_ops.py openpose arcface --brief 2>&1 | grep -E "^==" | cut -c1-170
