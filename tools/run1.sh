set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_run1_gpu.txt
python -m pytest tests -m gpu -x -q -s > gpurun_out/r2_run1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_run1_pytest.log
tail -5 gpurun_out/r2_run1_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_run1_bench.json 2> gpurun_out/r2_run1_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2_run1_bench.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_run1_bench_ref.json 2> gpurun_out/r2_run1_bench_ref.err; echo "ref rc=$?"
