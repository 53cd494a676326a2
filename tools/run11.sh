set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
R=r2_run11
python -m pytest tests/test_gpu_blend_adjoint_variants.py tests/test_gpu_parity.py tests/test_gpu_properties.py -m gpu -q -x > gpurun_out/${R}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${R}_pytest.log
tail -4 gpurun_out/${R}_pytest.log
python bench.py --steps 20 --warmup 5 --only-resident > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err; echo "bench rc=$?"; cat gpurun_out/${R}_bench.json | cut -c1-600
ncu --set full --clock-control none --import-source on -k "regex:blend_(backward|forward)" -s 2 -c 2 -o gpurun_out/${R}_blend python bench.py --steps 1 --warmup 1 --only-resident > gpurun_out/${R}_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
