set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
R=r2_run13
python bench.py --steps 20 --warmup 5 --only-resident > gpurun_out/${R}_bench_cfg2.json 2> gpurun_out/${R}_bench_cfg2.err; echo "bench rc=$?"; cat gpurun_out/${R}_bench_cfg2.json | cut -c1-600
python bench.py --steps 20 --warmup 5 --workload cfg3view --no-cpu-baseline > gpurun_out/${R}_bench_cfg3view.json 2> gpurun_out/${R}_bench_cfg3view.err; echo "bench rc=$?"; cat gpurun_out/${R}_bench_cfg3view.json | cut -c1-3000
tail -5 gpurun_out/${R}_bench_cfg3view.err
