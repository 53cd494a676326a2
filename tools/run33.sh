set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
R=r2_run34
(timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5) > gpurun_out/${R}_bench_n2.log 2>&1; echo "bench N=2 rc=$?"; grep "^{" gpurun_out/${R}_bench_n2.log | tail -1 | python -c "
import json,sys;d=json.loads(sys.stdin.read());print(d['value'], d['e2e']['value'], d['e2e']['mode'], d['e2e']['eager'])"
grep -i "capture\|error" gpurun_out/${R}_bench_n2.log | head -5 | cut -c1-300
