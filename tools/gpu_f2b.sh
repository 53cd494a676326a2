set -x
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_densify.py tests/test_gpu_fused.py -m gpu -q -x -s) > gpurun_out/pytest_f2.log 2>&1; echo "f2+fused rc=$?"; grep -E "^\[adam\]|^\[densify\]|passed|failed|Error|error" gpurun_out/pytest_f2.log | cut -c1-700 | tail -12
(timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:adam\|densify\|opacity_reset\|DeviceScan --csv --log-file gpurun_out/ncu_f2.csv python tools/f2_profile.py) > gpurun_out/ncu_f2.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_f2.log
