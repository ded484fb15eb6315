#!/bin/bash
# One GPU-box session: isolated test processes (a trapping kernel must not take
# the other suites down with it), each under its own timeout; logs land in
# gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
run() { # name, timeout, command...
  local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/session.log
  timeout "$t" "$@" > "gpurun_out/$name.log" 2>&1
  echo "exit $? ($name)" | tee -a gpurun_out/session.log
  tail -n 25 "gpurun_out/$name.log" | sed 's/^/    /' >> gpurun_out/session.log
}
: > gpurun_out/session.log
for step in "$@"; do
  case $step in
    post)    run post 900 python -m pytest tests/test_gpu_post.py -q -m gpu --tb=short ;;
    direct)  run ops_direct 900 python -m pytest tests/test_gpu_ops.py -q -m gpu --tb=short -k "not tcgen05" ;;
    tc)      run ops_tc 900 python -m pytest tests/test_gpu_ops.py -q -m gpu --tb=line -k "tcgen05 or stream_k" ;;
    nets_direct) run nets_direct 1500 python -m pytest tests/test_gpu_nets.py -q -m gpu --tb=short -k "direct" ;;
    nets)    run nets 1500 python -m pytest tests/test_gpu_nets.py -q -m gpu --tb=short -k "not direct" ;;
    all)     run all 2400 python -m pytest tests -x -q -m gpu ;;
    smoke)   run smoke 900 python __graft_entry__.py --smoke ;;
    bench)   run bench 1200 python bench.py ;;
    benchref) run benchref 1200 python bench.py --impl reference --steps 2 --warmup 1 ;;
    ncu_list) run ncu_list 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv \
                --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline ;;
    ncu_full) run ncu_full 1500 ncu --set full --clock-control none --import-source on -k regex:conv_tc \
                -s 100 -c 3 -f -o gpurun_out/prof_conv_tc python bench.py --steps 1 --warmup 1 --no-cpu-baseline ;;
    ncu_traffic) run ncu_traffic 1500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
                --clock-control none -k regex:conv_tc -s 279 -c 93 --csv --log-file gpurun_out/traffic.csv \
                python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-per-config ;;
    *)       run custom 1200 bash -c "$step" ;;
  esac
done
cat gpurun_out/session.log
