# e2e window per rank, spinning vs blocking waits (NG ranks)
python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-8} --master-addr 127.0.0.1 --master-port 29531 scripts/debug/e2e_ranks.py > gpurun_out/e2e_ranks_${NG:-8}.log 2>&1
tail -60 gpurun_out/e2e_ranks_${NG:-8}.log | cut -c1-260
