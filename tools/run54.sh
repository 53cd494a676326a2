set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
R=r2_run54
(time timeout 600 python -c "import __graft_entry__ as g; g.smoke()") > gpurun_out/${R}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${R}_smoke.log
(time timeout 1200 python -m pytest tests -x -q -m gpu) > gpurun_out/${R}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${R}_pytest.log
(time timeout 600 python bench.py) > gpurun_out/${R}_bench_default.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/${R}_bench_default.log | cut -c1-300
KR='regex:^(sh_forward|sh_backward|project_forward|project_backward|bin_|tile_s|blend_forward_kernel|blend_backward)'
timeout 600 ncu --set full --clock-control none --import-source on -k "$KR" -s 13 -c 13 -f -o gpurun_out/${R}_all python bench.py --steps 1 --warmup 1 --only-resident > gpurun_out/${R}_ncu.log 2>&1
ls -la gpurun_out/${R}_all.ncu-rep
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${R}_launches.log 2>&1
