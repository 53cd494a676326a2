# multi-GPU bench exactly as the driver launches it + the 2-GPU exchange test; usage: bash tools/gpu_multi.sh N
set -x
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_${N}.txt 2>&1
(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "exchange or multiview") > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_multi.log
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps 20 --warmup 5) > gpurun_out/bench_n${N}.log 2>&1; echo "bench rc=$?"; tail -2 gpurun_out/bench_n${N}.log
