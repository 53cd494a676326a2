set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
R=r2_run14
python bench.py --steps 20 --warmup 5 --workload cfg3view --no-cpu-baseline > gpurun_out/${R}_bench_cfg3view.json 2> gpurun_out/${R}_bench_cfg3view.err; echo "bench rc=$?"; cat gpurun_out/${R}_bench_cfg3view.json | cut -c1-1500
tail -5 gpurun_out/${R}_bench_cfg3view.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${R}_launches_cfg3view.csv python bench.py --steps 2 --warmup 1 --workload cfg3view --only-resident > gpurun_out/${R}_ncu.log 2>&1
