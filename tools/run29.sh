set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
R=r2_run29
for N in 8 4; do
  for tail in own nccl; do
  if [ $tail = nccl ]; then export GSR_NCCL_TAIL=1; else unset GSR_NCCL_TAIL; fi
  (timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N \
     bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline) > gpurun_out/${R}_bench_n${N}_$tail.log 2>&1; echo "bench N=$N $tail rc=$?"; grep "^{" gpurun_out/${R}_bench_n${N}_$tail.log | tail -1 | cut -c1-200
  done
done
