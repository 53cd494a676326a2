cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time timeout 28 python -m pytest tests/test_gpu_dropin_reference_callers.py -x -q -m gpu) > gpurun_out/r2_run59_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_run59_pytest.log
