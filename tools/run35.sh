set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
R=r2_run35
(time timeout 600 python -c "import __graft_entry__ as g; g.smoke()") > gpurun_out/${R}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${R}_smoke.log
(time timeout 2400 python -m pytest tests -x -q -m gpu) > gpurun_out/${R}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${R}_pytest.log
(time timeout 900 python bench.py) > gpurun_out/${R}_bench_default.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/${R}_bench_default.log | cut -c1-300
(time timeout 900 python bench.py --impl reference --steps 3 --warmup 1) > gpurun_out/${R}_bench_reference.log 2>&1; echo "ref rc=$?"; tail -1 gpurun_out/${R}_bench_reference.log | cut -c1-400
(time timeout 900 python bench.py --workload cfg4 --steps 10 --warmup 3 --no-cpu-baseline) > gpurun_out/${R}_bench_cfg4.log 2>&1; echo "cfg4 rc=$?"; tail -1 gpurun_out/${R}_bench_cfg4.log | cut -c1-300
