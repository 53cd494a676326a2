set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s > gpurun_out/r2_run5_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_run5_pytest.log
tail -8 gpurun_out/r2_run5_pytest.log
for mode in pixel scan scan4; do
GSR_BWD_KERNEL=$mode python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_run5_bench_$mode.json 2> gpurun_out/r2_run5_bench_$mode.err; echo "bench $mode rc=$?"
done
python - <<'PY'
import json
for m in ("pixel","scan","scan4"):
    try:
        d=json.load(open(f"gpurun_out/r2_run5_bench_{m}.json")); print(m, d["value"], d["e2e"]["value"], d["gpu_launches"], d["stages_ms"]); print(d["ref_cuda_ext"].get("speedup"), d.get("cuda_graph"))
    except Exception as e: print(m, "ERR", e)
PY
ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 100 --csv --log-file gpurun_out/r2_run5_launches.csv python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2_run5_ncu_launch.log 2>&1
