set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
R=r2_run30
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${R}_bench.err; python -c "
import json;d=json.loads([l for l in open('gpurun_out/${R}_bench.json') if l.startswith('{')][-1]);print(d['value'], d['ms_per_step'], d['stages_ms'], d['fused_operator']['value'], d['e2e']['value'], d['cuda_graph'])"
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload cfg3view > gpurun_out/${R}_bench_cfg3view.json 2> gpurun_out/${R}_bench_cfg3view.err; echo "bench rc=$?"; python -c "
import json;d=json.loads([l for l in open('gpurun_out/${R}_bench_cfg3view.json') if l.startswith('{')][-1]);print(d['value'], d['ms_per_step'], d['stages_ms'], d['fused_operator']['value'], d['e2e']['value'], d['cuda_graph']['value'], d['ref_cuda_ext']['speedup'])"
