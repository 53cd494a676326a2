# one optimisation iteration on the GPU box: parity tests, side-by-side timing vs the reference ext, bench, launch list
set -x
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_perf_vs_ref.py) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
(timeout 900 python -m pytest tests/test_gpu_perf_vs_ref.py -m gpu -q -s) > gpurun_out/perf_vs_ref.log 2>&1; echo "perf rc=$?"; grep -E "ms_per_view|speedup|rasterize_|sort_|map_|count" gpurun_out/perf_vs_ref.log
(timeout 900 python bench.py --steps 20 --warmup 5) > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -2 gpurun_out/bench.log
