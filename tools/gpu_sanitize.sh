# compute-sanitizer passes over the small-shape parity tests (memcheck: OOB / misaligned; racecheck: shared-memory hazards;
# synccheck: divergent barriers).  Slow: only the tiny scenes.
set -x
mkdir -p gpurun_out
SEL='tiny_17 or deg0_700 or block_widths or nd_channels or tight_binning_degenerate or fast_binning_large or multiview or nd_colors'
SEL2='1k_64x64_deg0 or 2k_96x64_deg4 or empty_and_no_depth or 11-11 or 37-53 or 64-48 or errors_and_determinism'
for tool in memcheck racecheck synccheck; do
  (timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 \
      python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL") > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool parity rc=$?"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" gpurun_out/sanitize_$tool.log | tail -3
  (timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 \
      python -m pytest tests/test_gpu_fused.py tests/test_gpu_loss.py -m gpu -q -x -k "$SEL2") > gpurun_out/sanitize_${tool}_fused_loss.log 2>&1
  echo "$tool fused+loss rc=$?"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" gpurun_out/sanitize_${tool}_fused_loss.log | tail -3
done
