set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
R=r2_run43
(timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 \
     bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline) > gpurun_out/${R}_bench_n8.log 2>&1; echo "bench N=8 rc=$?"
grep "^{" gpurun_out/${R}_bench_n8.log | tail -1 | python -c "
import json,sys;d=json.loads(sys.stdin.read());st=d['stages_ms'];print(d['value'], d['ms_per_step'], {k[:30]:v for k,v in st.items() if 'exchange' in k}, d['exchange_check'], d['e2e']['value'], d['e2e']['mode'], d['e2e']['eager'])"
