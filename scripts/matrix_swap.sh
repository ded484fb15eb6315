for dbg in 32 34; do echo "== patch variant TRB_TC_DEBUG=$dbg"; TRB_TC_SWAP=1 TRB_TC_DEBUG=$dbg python scripts/bench_conv.py --short "7x7 128->128" | tail -3; done
for dbg in 32 35; do echo "== normal TRB_TC_DEBUG=$dbg"; TRB_TC_SWAP=0 TRB_TC_DEBUG=$dbg python scripts/bench_conv.py --short "7x7 128->128" "vgg 3x3 512" | grep -v producer | tail -4; done
for sw in 1 0; do echo "== TRB_TC_SWAP=$sw"; TRB_TC_SWAP=$sw python scripts/bench_conv.py; done
for sub in 1 3; do echo "== patch variant SUB=$sub"; TRB_TC_SWAP=1 TRB_TC_SUB=$sub python scripts/bench_conv.py "7x7 128->128"; done
