set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
R=r2_run12
python -m pytest tests/test_gpu_blend_adjoint_variants.py tests/test_gpu_parity.py tests/test_gpu_properties.py -m gpu -q -x -s > gpurun_out/${R}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${R}_pytest.log
grep "adjoint variants\] tr " gpurun_out/${R}_pytest.log; tail -4 gpurun_out/${R}_pytest.log
for mode in tr tr8 tr32; do
GSR_BWD_KERNEL=$mode python bench.py --steps 20 --warmup 5 --only-resident > gpurun_out/${R}_bench_${mode}.json 2> gpurun_out/${R}_bench_${mode}.err; echo "bench $mode rc=$?"; cat gpurun_out/${R}_bench_${mode}.json | cut -c1-600
done
