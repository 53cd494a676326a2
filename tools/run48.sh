set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
R=r2_run48
(time timeout 600 python -m pytest tests/test_gpu_binning_device.py tests/test_gpu_parity.py tests/test_gpu_properties.py -x -q -m gpu) > gpurun_out/${R}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${R}_pytest.log
timeout 300 python bench.py --steps 30 --warmup 5 --only-resident > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err; echo "bench rc=$?"; tail -1 gpurun_out/${R}_bench.json | cut -c1-600
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 2 --warmup 1 --only-resident > gpurun_out/${R}_launches.log 2>&1
timeout 400 python bench.py --steps 10 --warmup 3 --only-resident --workload cfg4 > gpurun_out/${R}_bench_cfg4.json 2> gpurun_out/${R}_bench_cfg4.err; echo "cfg4 rc=$?"; tail -1 gpurun_out/${R}_bench_cfg4.json | cut -c1-600
