set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
R=r2_run56
(time timeout 300 python -c "import __graft_entry__ as g; g.smoke()") > gpurun_out/${R}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${R}_smoke.log
(time timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fused.py -x -q -m gpu) > gpurun_out/${R}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${R}_pytest.log
GSR_NVTX=1 timeout 200 ncu --nvtx --print-nvtx-rename kernel --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/${R}_nvtx_launches.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_nvtx.log 2>&1; echo "nvtx rc=$?"
