set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
R=r2_run20
(time timeout 900 python -m pytest tests/test_gpu_train_view_parallel.py tests/test_gpu_bucket_odd_n.py "tests/test_gpu_parity.py::test_exchange_gradients_matches_plain_allreduce_2gpu" -m gpu -q -x -s) > gpurun_out/${R}_pytest_2gpu.log 2>&1; echo "2gpu rc=$?"; grep -E "^\[|passed|failed|Error|error" gpurun_out/${R}_pytest_2gpu.log | cut -c1-400 | tail -14
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5) > gpurun_out/${R}_bench_n2.log 2>&1; echo "bench N=2 rc=$?"; grep "^{" gpurun_out/${R}_bench_n2.log | tail -1 | cut -c1-300
(timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline) > gpurun_out/${R}_bench_n1.log 2>&1; echo "bench N=1 rc=$?"; grep "^{" gpurun_out/${R}_bench_n1.log | tail -1 | cut -c1-300
