#!/bin/bash
# compute-sanitizer over a subset of the GPU parity tests (the full-size cases are left out for run time).
# Usage on the GPU box:  bash tools/sanitize.sh      (summaries under gpurun_out/sanitizer_*.log)
mkdir -p gpurun_out
SEL='test_op or test_small_mining_match_known or bucket_overflow or test_encode_inside_mask or fused_encode_bucket or test_nms_bboxes_vs_oracle or nms_dense or parse_by_class_variants or parse_by_class_golden or test_dual or hard_negative or routing or handoff or device_gather or logit_gaps or vote or host_resident or (layout_hint and size0) or peer_exchange'
for tool in memcheck initcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests -m gpu -q -x -k "$SEL" -p no:cacheprovider > gpurun_out/sanitizer_$tool.full 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|error" gpurun_out/sanitizer_$tool.full | tail -5 > gpurun_out/sanitizer_$tool.log
  grep -E "^=========     (Race|Invalid|Uninitialized|Potential|at |by )" gpurun_out/sanitizer_$tool.full | sort | uniq -c | sort -rn | head -30 >> gpurun_out/sanitizer_$tool.log
  echo "== $tool"; cat gpurun_out/sanitizer_$tool.log | head -12
done
