set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
R=r2_run36
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${R}_launches_cfg4.csv python bench.py --steps 2 --warmup 1 --workload cfg4 --only-resident > gpurun_out/${R}_ncu.log 2>&1
