# session U: FC patch-buffer depth, halo second issuer on the 64-filter layers, ncu captures
timeout 600 python -m pytest tests/test_gpu_conv_patch.py tests/test_gpu_capi.py -m gpu -q -x 2>&1 | grep -E "^E  |passed|failed" | head -12 | cut -c1-300
python scripts/bench_patch.py fc 2>&1 | tail -20
python scripts/profile_ops.py arcface --brief 2>&1 | grep -E "^==|25088|tcgen05"
for h in 0 1; do echo "== TRB_TC_ISSUERS_HALO=$h"; TRB_TC_ISSUERS_HALO=$h python scripts/profile_ops.py openpose arcface 2>&1 | grep -E "^==|64->  64|64-> 128|tcgen05"; done
# ncu: every metric of the first conv_tc launches (VGG conv1_2 = 64->64 halo) and the patch kernels
ncu --set full --clock-control none -k regex:conv_tc -c 12 --csv --page raw --log-file gpurun_out/r2u_ncu_conv_tc.csv python scripts/profile_ops.py openpose --brief > gpurun_out/r2u_ncu1.log 2>&1
ncu --set full --clock-control none -k regex:conv_patch -c 45 --csv --page raw --log-file gpurun_out/r2u_ncu_conv_patch.csv python scripts/profile_ops.py openpose --brief > gpurun_out/r2u_ncu2.log 2>&1
ncu --set full --clock-control none -k regex:"pose_|resize_u8|det_|l2_norm|stem_mma|maxpool|export_" -c 24 --csv --page raw --log-file gpurun_out/r2u_ncu_aux.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-per-config > gpurun_out/r2u_ncu3.log 2>&1
ls -la gpurun_out/r2u_*
