for sub in 2 3; do echo "== TRB_TC_CTAS=0 TRB_TC_SUB=$sub"; TRB_TC_CTAS=0 TRB_TC_SUB=$sub python scripts/bench_conv.py "7x7" "arcface 3x3 256" "vgg 3x3 256" "vgg 3x3 512"; done
