set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
R=r2_run40
(time timeout 300 python -m pytest "tests/test_gpu_parity.py::test_exchange_gradients_matches_plain_allreduce_2gpu" -m gpu -q -x -s) > gpurun_out/${R}_pytest_2gpu.log 2>&1; echo "2gpu rc=$?"; grep -E "^\[|passed|failed|Error|error" gpurun_out/${R}_pytest_2gpu.log | cut -c1-400 | tail -8
for t in push nccl; do
if [ $t = push ]; then export GSR_OWN_TAIL=push; else unset GSR_OWN_TAIL; fi
(timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5) > gpurun_out/${R}_bench_n2_$t.log 2>&1; echo "bench N=2 $t rc=$?"
grep "^{" gpurun_out/${R}_bench_n2_$t.log | tail -1 | python -c "
import json,sys;d=json.loads(sys.stdin.read());st=d['stages_ms'];print(d['value'], d['ms_per_step'], {k[:30]:v for k,v in st.items() if 'exchange' in k}, d['exchange_check'])"
done
tail -5 gpurun_out/${R}_bench_n2_push.log | cut -c1-300
