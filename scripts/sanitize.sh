#!/bin/bash
# compute-sanitizer over the single-op GPU test matrix (the reference has no
# sanitizer runs; every kernel here is hand-written, so memcheck + racecheck are part of the
# parity bar).  Run on the GPU box:   bash scripts/sanitize.sh [out.log]
# memcheck: out-of-bounds / misaligned global + shared accesses of every kernel variant
# (tcgen05 plain / halo / stream-K / two issuers, resident-patch kernel with and without
# stream-K, mma.sync, direct, post-processing); racecheck: shared-memory hazards of the
# hand-rolled mbarrier pipelines.  The summaries are kept under profiles/.
OUT=${1:-gpurun_out/sanitizer.log}
mkdir -p "$(dirname "$OUT")"
: > "$OUT"
SAN=/usr/local/cuda/bin/compute-sanitizer
run() {   # name, tool, env assignments..., then "--", then pytest selection
  local name=$1 tool=$2; shift 2
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  echo "=== $name ($tool ${envs[*]})" | tee -a "$OUT"
  env "${envs[@]}" timeout 1500 $SAN --tool "$tool" --error-exitcode 3 --print-limit 5 \
      python -m pytest -m gpu -x -q "$@" 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Invalid|Hazard|error" | tail -6 | tee -a "$OUT"
}
PATCH="tests/test_gpu_conv_patch.py"
OPS="tests/test_gpu_ops.py"
run patch-default   memcheck  -- $PATCH -k "matches_reference or epilogues or stream_k"
run patch-sk-always memcheck  TRB_PT_SK=2 -- $PATCH -k "matches_reference or tile_geometries"
run patch-generic   memcheck  TRB_PT_GENERIC=1 -- $PATCH -k "matches_reference or epilogues"
run patch-race      racecheck -- $PATCH -k "matches_reference or stream_k_epilogues"
run tc-default      memcheck  -- $OPS -k "conv or sepconv"
run tc-fat-kernel   memcheck  TRB_TC_LEAN=0 -- $OPS -k "conv_matches_reference"
run tc-resident-2   memcheck  TRB_TC_RESIDENT=2 -- $OPS -k "conv_matches_reference"
run tc-resident-off memcheck  TRB_TC_RESIDENT=0 -- $OPS -k "conv_matches_reference"
run tc-issuers-off  memcheck  TRB_TC_ISSUERS=0 -- $OPS -k "conv_matches_reference or conv_stream_k"
run tc-sk-always    memcheck  TRB_TC_SK=2 -- $OPS -k "conv_matches_reference or conv_stream_k"
run tc-halo-off     memcheck  TRB_TC_HALO=0 -- $OPS -k "conv_matches_reference or conv_stream_k"
run tc-race         racecheck -- $OPS -k "conv_matches_reference or conv_stream_k"
run post            memcheck  -- tests/test_gpu_post.py
run post-race       racecheck -- tests/test_gpu_post.py -k "decode or parse"
echo "done" | tee -a "$OUT"
