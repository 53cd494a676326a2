# the driver's scaling run at N = 8 (and 2, 4 on the same box): torchrun + bench.py, one JSON line each
set -x
mkdir -p gpurun_out
for N in 8 4 2; do
  (timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N \
     bench.py --gpus $N --steps 20 --warmup 5) > gpurun_out/bench_n${N}.log 2>&1; echo "bench N=$N rc=$?"; grep "^{" gpurun_out/bench_n${N}.log | tail -1 | cut -c1-400
done
