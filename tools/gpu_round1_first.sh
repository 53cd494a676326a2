set -x
mkdir -p gpurun_out
nvidia-smi -L | head -2; nproc; free -g | head -2
(timeout 600 python __graft_entry__.py smoke) > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -15 gpurun_out/smoke.log
(timeout 900 python tests/golden/gen_golden_ref_cuda.py) > gpurun_out/gen_golden.log 2>&1; echo "golden rc=$?"; tail -5 gpurun_out/gen_golden.log
(timeout 1500 python -m pytest tests -m gpu -q -s -x --deselect tests/test_gpu_perf_vs_ref.py) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest_gpu.log
(timeout 900 python -m pytest tests/test_gpu_perf_vs_ref.py -m gpu -q -s) > gpurun_out/perf_vs_ref.log 2>&1; echo "perf rc=$?"; tail -45 gpurun_out/perf_vs_ref.log
(timeout 900 python bench.py --steps 10 --warmup 3) > gpurun_out/bench1.log 2>&1; echo "bench rc=$?"; tail -5 gpurun_out/bench1.log
