set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
R=r2_run45
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "oracle or golden or sh" > gpurun_out/${R}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${R}_pytest.log
for t in 1 0; do
GSR_SH_TMA=$t ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sh_forward -c 6 --csv --log-file gpurun_out/${R}_sh_$t.csv python bench.py --steps 3 --warmup 2 --only-resident > gpurun_out/${R}_ncu_$t.log 2>&1
grep "sh_forward" gpurun_out/${R}_sh_$t.csv | awk -F'","' '{print $5, $(NF)}' | cut -c1-60,200-260 | tail -3
GSR_SH_TMA=$t python bench.py --steps 30 --warmup 5 --only-resident 2>/dev/null | cut -c1-200
done
