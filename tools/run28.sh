set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
R=r2_run28
(time timeout 900 python -m pytest tests/test_gpu_train_view_parallel.py tests/test_gpu_bucket_odd_n.py "tests/test_gpu_parity.py::test_exchange_gradients_matches_plain_allreduce_2gpu" "tests/test_gpu_parity.py::test_sh_backward_multiview_equals_sum_of_views" -m gpu -q -x -s) > gpurun_out/${R}_pytest_2gpu.log 2>&1; echo "2gpu rc=$?"; grep -E "^\[|passed|failed|Error|error" gpurun_out/${R}_pytest_2gpu.log | cut -c1-400 | tail -14
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5) > gpurun_out/${R}_bench_n2.log 2>&1; echo "bench N=2 rc=$?"; grep "^{" gpurun_out/${R}_bench_n2.log | tail -1 | cut -c1-200
(GSR_NCCL_TAIL=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5) > gpurun_out/${R}_bench_n2_nccl.log 2>&1; echo "bench N=2 nccl rc=$?"; grep "^{" gpurun_out/${R}_bench_n2_nccl.log | tail -1 | cut -c1-200
tail -5 gpurun_out/${R}_bench_n2.log | cut -c1-300
