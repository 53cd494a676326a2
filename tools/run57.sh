set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
R=r2_run57
(time timeout 45 python -m pytest "tests/test_gpu_parity.py::test_exchange_gradients_matches_plain_allreduce_2gpu" tests/test_gpu_bucket_odd_n.py -m gpu -q -x) > gpurun_out/${R}_pytest_2gpu.log 2>&1; echo "2gpu rc=$?"; tail -2 gpurun_out/${R}_pytest_2gpu.log
(timeout 50 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 15 --warmup 4) > gpurun_out/${R}_bench_n2.log 2>&1; echo "bench N=2 rc=$?"
grep "^{" gpurun_out/${R}_bench_n2.log | tail -1 | cut -c1-400
