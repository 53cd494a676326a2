set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s > gpurun_out/r2_run3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_run3_pytest.log
tail -15 gpurun_out/r2_run3_pytest.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_run3_bench.json 2> gpurun_out/r2_run3_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r2_run3_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_run3_bench.json")); print(d["value"], d["e2e"]["value"], d["gpu_launches"], d["stages_ms"]); print(d["ref_cuda_ext"].get("speedup"))
PY
# launch list of the bench command (shares) and full captures of the two adjoint kernels
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 160 --csv --log-file gpurun_out/r2_run3_launches.csv python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2_run3_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:blend_backward_kernel -s 2 -c 1 -o gpurun_out/r2_run3_bwd_pixel python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2_run3_ncu1.log 2>&1
GSR_BWD_KERNEL=scan ncu --set full --clock-control none --import-source on -k regex:blend_backward_scan -s 2 -c 1 -o gpurun_out/r2_run3_bwd_scan python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2_run3_ncu2.log 2>&1
ls -la gpurun_out | tail -8
