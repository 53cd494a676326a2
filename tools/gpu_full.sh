# what the driver does at round end: smoke, the whole -m gpu suite, the default bench, the reference arm
set -x
mkdir -p gpurun_out
(time timeout 600 python -c "import __graft_entry__ as g; g.smoke()") > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
(time timeout 2400 python -m pytest tests -x -q -m gpu) > gpurun_out/pytest_gpu_full.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu_full.log
(time timeout 900 python bench.py) > gpurun_out/bench_default.log 2>&1; echo "bench rc=$?"; tail -4 gpurun_out/bench_default.log | cut -c1-600
(time timeout 900 python bench.py --impl reference --steps 5 --warmup 1) > gpurun_out/bench_reference.log 2>&1; echo "ref rc=$?"; tail -4 gpurun_out/bench_reference.log | cut -c1-900
