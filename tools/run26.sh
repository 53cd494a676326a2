set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
R=r2_run26
for N in 8 4; do
  (timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N \
     bench.py --gpus $N --steps 20 --warmup 5) > gpurun_out/${R}_bench_n${N}.log 2>&1; echo "bench N=$N rc=$?"; grep "^{" gpurun_out/${R}_bench_n${N}.log | tail -1 | cut -c1-300
done
(timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline) > gpurun_out/${R}_bench_n1.log 2>&1; echo "bench N=1 rc=$?"
