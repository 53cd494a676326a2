set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
R=r2_run19
nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max --format=csv
python - <<'PY'
import torch,time
x=torch.empty(64*1024*1024//4,dtype=torch.float32).pin_memory(); d=torch.empty_like(x,device='cuda')
for _ in range(3): d.copy_(x,non_blocking=True)
torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
e0.record(); 
for _ in range(10): d.copy_(x,non_blocking=True)
e1.record(); torch.cuda.synchronize(); print('H2D GB/s', 10*64/1024/ (e0.elapsed_time(e1)*1e-3))
e0.record(); 
for _ in range(10): x.copy_(d,non_blocking=True)
e1.record(); torch.cuda.synchronize(); print('D2H GB/s', 10*64/1024/ (e0.elapsed_time(e1)*1e-3))
PY
(time python bench.py --no-cpu-baseline) > gpurun_out/${R}_bench_default.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/${R}_bench_default.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['e2e']['ms_per_step'])"
