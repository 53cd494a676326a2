set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
R=r2_run16
(time timeout 2400 python -m pytest tests -x -q -m gpu) > gpurun_out/${R}_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/${R}_pytest.log
python bench.py --steps 20 --warmup 5 --only-resident > gpurun_out/${R}_bench_cfg2.json 2> gpurun_out/${R}_bench_cfg2.err; echo "bench rc=$?"; cat gpurun_out/${R}_bench_cfg2.json | cut -c1-600
python bench.py --steps 20 --warmup 5 --workload cfg3view --only-resident > gpurun_out/${R}_bench_cfg3view.json 2> gpurun_out/${R}_bench_cfg3view.err; echo "bench rc=$?"; cat gpurun_out/${R}_bench_cfg3view.json | cut -c1-1500
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${R}_launches_cfg3view.csv python bench.py --steps 2 --warmup 1 --workload cfg3view --only-resident > gpurun_out/${R}_ncu.log 2>&1
