set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
R=r2_run18
(time python bench.py) > gpurun_out/${R}_bench_default.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/${R}_bench_default.log | cut -c1-400
KR='regex:^(sh_forward|sh_backward|project_forward|project_backward|bin_|tile_s|blend_forward_kernel|blend_backward)'
ncu --set full --clock-control none --import-source on -k "$KR" -s 14 -c 14 -o gpurun_out/${R}_all python bench.py --steps 1 --warmup 1 --only-resident > gpurun_out/${R}_ncu.log 2>&1
ls -la gpurun_out/${R}_all.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/${R}_launches.log 2>&1
