#!/bin/bash
# compute-sanitizer over the small-shape GPU tests (race detection / sanitizers).
#   memcheck  : out-of-bounds / misaligned global, shared and TMA accesses
#   racecheck : shared-memory hazards inside a CTA (the staged epilogues, the per-warp top-k lists, block reductions)
#   synccheck : divergent / mismatched barriers (named barriers of the epilogue warps, cluster barriers)
# The sanitizer slows kernels by 10-100x: only the tests whose shapes are small are selected, each tool under its own timeout.
# (It does not see races BETWEEN kernels on different streams: tests/test_schedule_races.py covers those on the CPU.)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
SEL='test_ctx_and_errors or test_philox_table_bit_exact or test_corrupt_fwd_and_dense_masks or test_mul_mask or test_mse_loss_fwd_bwd or test_mixed_loss_and_monitor or test_adam_device_step_counter or (test_linear_tcgen05_engine and (128-64-64 or 128-256-128 or 200-192-192 or 300-328-72)) or (test_linear_f32_engine) or (test_tma_store_epilogue_is_bit_identical and (128-192-328 or 100-832-128))'
for TOOL in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $TOOL"
  timeout -s KILL 420 compute-sanitizer --tool $TOOL --error-exitcode 9 --print-limit 20 \
      python -m pytest tests/test_gpu_kernels.py tests/test_gpu_wgrad_sqnorm.py -q -m gpu -p no:cacheprovider -x -k "$SEL" \
      > gpurun_out/sanitize_$TOOL.log 2>&1
  echo "rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Error|RACECHECK SUMMARY" gpurun_out/sanitize_$TOOL.log | tail -6
done
echo "== golden training steps (fused path, small nets) under memcheck"
timeout -s KILL 420 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_gpu_training.py -q -m gpu -p no:cacheprovider -x -k "test_fused_step_matches_reference and (emb_small or emb_k2)" \
    > gpurun_out/sanitize_training.log 2>&1
echo "rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_training.log | tail -4
echo "== done"
