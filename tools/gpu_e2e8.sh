# 8-GPU e2e diagnosis: default (NUMA binding + NVLink peer SH exchange in the autograd leg), without binding, without copies
set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_8.txt 2>&1
run() { tag=$1; shift; (env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
     bench.py --gpus 8 --steps 12 --warmup 4) > gpurun_out/bench_n8_$tag.log 2>&1; echo "$tag rc=$?"; grep "^{" gpurun_out/bench_n8_$tag.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$tag', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), d['e2e']['ms_per_step'], d['config'].get('host_numa_binding'), d['config'].get('sh_gradient_exchange'))"; }
run default A=1
run nonuma GSR_NUMA_BIND=0
run nocopy GSR_E2E_NOCOPY=1
