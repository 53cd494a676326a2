set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
R=r2_run38
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "needles" -s > gpurun_out/${R}_pytest.log 2>&1; echo "pytest rc=$?"; grep -v "^\[parity\]" gpurun_out/${R}_pytest.log | tail -30 | cut -c1-250
