set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
R=r2_run37
python -m pytest tests/test_gpu_binning_device.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/${R}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${R}_pytest.log
python bench.py --steps 20 --warmup 5 --only-resident > gpurun_out/${R}_bench_cfg2.json 2> gpurun_out/${R}_bench_cfg2.err; echo "bench rc=$?"; cat gpurun_out/${R}_bench_cfg2.json | cut -c1-500
python bench.py --steps 10 --warmup 3 --only-resident --workload cfg4 > gpurun_out/${R}_bench_cfg4.json 2> gpurun_out/${R}_bench_cfg4.err; echo "bench rc=$?"; cat gpurun_out/${R}_bench_cfg4.json | cut -c1-500
python bench.py --steps 20 --warmup 5 --only-resident --workload cfg3view > gpurun_out/${R}_bench_cfg3view.json 2> gpurun_out/${R}_bench_cfg3view.err; echo "bench rc=$?"; cat gpurun_out/${R}_bench_cfg3view.json | cut -c1-500
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${R}_launches_cfg4.csv python bench.py --steps 2 --warmup 1 --workload cfg4 --only-resident > gpurun_out/${R}_ncu.log 2>&1
