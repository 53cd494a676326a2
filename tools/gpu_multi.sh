# multi-GPU bench exactly as the driver launches it; usage: bash tools/gpu_multi.sh N
set -x
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_${N}.txt 2>&1
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps 20 --warmup 5) > gpurun_out/bench_n${N}.log 2>&1; echo "bench rc=$?"; tail -3 gpurun_out/bench_n${N}.log
