python scripts/bench_patch.py issue > gpurun_out/r2e_issue.log 2>&1; cat gpurun_out/r2e_issue.log
for v in 0 1 2; do echo "== TRB_PATCH=$v"; TRB_PATCH=$v python scripts/profile_ops.py openpose --brief 2>&1 | grep -E "^==|k7|tcgen05|k3 s1  (128|256|512)"; done > gpurun_out/r2e_net.log 2>&1; cat gpurun_out/r2e_net.log
