set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests/test_gpu_blend_units.py tests/test_gpu_loss.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2_run8_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_run8_pytest.log
tail -4 gpurun_out/r2_run8_pytest.log
for mode in pixel scan scanp; do
GSR_BWD_KERNEL=$mode python bench.py --steps 20 --warmup 5 --only-resident > gpurun_out/r2_run8_bench_$mode.json 2> gpurun_out/r2_run8_bench_$mode.err; echo "bench $mode rc=$?"; cat gpurun_out/r2_run8_bench_$mode.json | cut -c1-600
done
GSR_BWD_KERNEL=scanp python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "oracle or golden" > gpurun_out/r2_run8_pytest_scanp.log 2>&1; tail -2 gpurun_out/r2_run8_pytest_scanp.log
KR='regex:^(sh_forward|sh_backward|project_forward|project_backward|bin_|tile_s|blend_forward_kernel|blend_backward)'
ncu --set full --clock-control none --import-source on -k "$KR" -s 13 -c 13 -o gpurun_out/r2_run8_all python bench.py --steps 1 --warmup 1 --only-resident > gpurun_out/r2_run8_ncu2.log 2>&1
ls -la gpurun_out/*.ncu-rep
