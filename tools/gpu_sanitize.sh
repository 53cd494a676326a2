# compute-sanitizer passes over the small-shape parity tests (memcheck: OOB / misaligned; racecheck: shared-memory hazards;
# synccheck: divergent barriers).  Slow: only the tiny scenes.
set -x
mkdir -p gpurun_out
SEL='tiny_17 or deg0_700 or block_widths or nd_channels or tight_binning_degenerate or fast_binning_large'
for tool in memcheck racecheck synccheck; do
  (timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 \
      python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL") > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" gpurun_out/sanitize_$tool.log | tail -3
done
