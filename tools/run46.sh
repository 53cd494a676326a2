set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests/test_gpu_blend_adjoint_variants.py -m gpu -q -x -k tma > gpurun_out/r2_run46_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_run46_pytest.log
