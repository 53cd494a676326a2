set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
R=r2_run42
(time timeout 300 python -m pytest "tests/test_gpu_parity.py::test_exchange_gradients_matches_plain_allreduce_2gpu" "tests/test_gpu_parity.py::test_sh_backward_multiview_equals_sum_of_views" -m gpu -q -x) > gpurun_out/${R}_pytest_2gpu.log 2>&1; echo "2gpu rc=$?"; tail -2 gpurun_out/${R}_pytest_2gpu.log
(timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5) > gpurun_out/${R}_bench_n2.log 2>&1; echo "bench N=2 rc=$?"
grep "^{" gpurun_out/${R}_bench_n2.log | tail -1 | python -c "
import json,sys;d=json.loads(sys.stdin.read());st=d['stages_ms'];print(d['value'], d['ms_per_step'], {k[:30]:v for k,v in st.items() if 'exchange' in k}, d['exchange_check'], d['e2e']['value'])"
