timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | cut -c1-300
python scripts/profile_ops.py arcface --brief 2>&1 | grep -E "^==|conv \*|stem|tcgen05"
