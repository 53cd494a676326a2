set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
R=r2_run55
(time timeout 400 python bench.py --workload cfg4 --steps 10 --warmup 3 --no-cpu-baseline) > gpurun_out/${R}_bench_cfg4.log 2>&1; echo "cfg4 rc=$?"; tail -1 gpurun_out/${R}_bench_cfg4.log | cut -c1-200
(time timeout 300 python bench.py --workload cfg3view --steps 30 --warmup 5 --no-cpu-baseline) > gpurun_out/${R}_bench_cfg3view.log 2>&1; echo "cfg3view rc=$?"; tail -1 gpurun_out/${R}_bench_cfg3view.log | cut -c1-200
