cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time timeout 50 python -m pytest tests/test_gpu_loss.py tests/test_gpu_densify.py tests/test_gpu_binning_device.py -x -q -m gpu) > gpurun_out/r2_run58_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_run58_pytest.log
