set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
R=r2_run47
(time timeout 1200 python -m pytest tests -x -q -m gpu) > gpurun_out/${R}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${R}_pytest.log
for m in 1 0 1 0; do
  GSR_BLOCK_MASK=$m timeout 300 python bench.py --steps 30 --warmup 5 --only-resident > gpurun_out/${R}_bench_mask$m.json 2> gpurun_out/${R}_bench_mask$m.err; echo "bench mask=$m rc=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/${R}_bench_mask$m.json").read().strip().splitlines()[-1])
print("mask=$m", round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["stages_ms"].items()})
PY
done
timeout 300 python bench.py --steps 20 --warmup 5 --only-resident --workload cfg3view > gpurun_out/${R}_bench_cfg3view.json 2>&1; tail -1 gpurun_out/${R}_bench_cfg3view.json | cut -c1-900
