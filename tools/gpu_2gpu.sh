set -x
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_train_view_parallel.py "tests/test_gpu_parity.py::test_exchange_gradients_matches_plain_allreduce_2gpu" -m gpu -q -x -s) > gpurun_out/pytest_2gpu.log 2>&1; echo "2gpu rc=$?"; grep -E "^\[|passed|failed|Error|error" gpurun_out/pytest_2gpu.log | cut -c1-400 | tail -14
