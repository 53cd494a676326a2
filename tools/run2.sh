set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s > gpurun_out/r2_run2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_run2_pytest.log
tail -5 gpurun_out/r2_run2_pytest.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_run2_bench_pixel.json 2> gpurun_out/r2_run2_bench_pixel.err; echo "bench rc=$?"
GSR_BWD_KERNEL=scan python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_run2_bench_scan.json 2> gpurun_out/r2_run2_bench_scan.err; echo "bench rc=$?"
python - <<'PY'
import json
for k in ("pixel","scan"):
    try:
        d=json.load(open(f"gpurun_out/r2_run2_bench_{k}.json")); print(k, d["value"], d["stages_ms"])
    except Exception as e: print(k, "ERR", e)
PY
