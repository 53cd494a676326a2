# f2 row on the GPU box: parity tests of the optimizer / densification kernels, the cfg3 training harness, an ncu pass
# over the f2 kernels (time + DRAM bytes per launch), and a short bench to confirm the hot path is unchanged
set -x
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_densify.py -m gpu -q -x -s) > gpurun_out/pytest_f2.log 2>&1; echo "f2 rc=$?"; grep -E "^\[adam\]|^\[densify\]|passed|failed|Error|error" gpurun_out/pytest_f2.log | cut -c1-700 | tail -12
(time timeout 900 python -m pytest tests/test_gpu_train_densify.py -m gpu -q -x -s) > gpurun_out/pytest_cfg3.log 2>&1; echo "cfg3 rc=$?"; grep -E "^\[cfg3\]|passed|failed|Error|error" gpurun_out/pytest_cfg3.log | cut -c1-700 | tail -12
(timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:adam\|densify\|opacity_reset\|DeviceScan --csv --log-file gpurun_out/ncu_f2.csv python tools/f2_profile.py) > gpurun_out/ncu_f2.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_f2.log
(timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline) > gpurun_out/bench_f2.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_f2.log | cut -c1-1500
