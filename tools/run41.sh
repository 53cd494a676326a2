set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
R=r2_run41
export GSR_OWN_TAIL=push
for N in 8 4; do
  (timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N \
     bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline) > gpurun_out/${R}_bench_n${N}_push.log 2>&1; echo "bench N=$N push rc=$?"
  grep "^{" gpurun_out/${R}_bench_n${N}_push.log | tail -1 | python -c "
import json,sys;d=json.loads(sys.stdin.read());st=d['stages_ms'];print(d['value'], d['ms_per_step'], {k[:30]:v for k,v in st.items() if 'exchange' in k}, d['exchange_check'], d['e2e']['value'])"
done
