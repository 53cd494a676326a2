set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s > gpurun_out/r2_run4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_run4_pytest.log
tail -12 gpurun_out/r2_run4_pytest.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_run4_bench.json 2> gpurun_out/r2_run4_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r2_run4_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_run4_bench.json")); print(d["value"], d["e2e"]["value"], d["gpu_launches"], d["stages_ms"]); print(d["ref_cuda_ext"].get("speedup"))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 100 --csv --log-file gpurun_out/r2_run4_launches.csv python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2_run4_ncu_launch.log 2>&1
