set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests/test_gpu_blend_units.py tests/test_gpu_loss.py tests/test_gpu_fused.py tests/test_gpu_densify.py -m gpu -q -x -s > gpurun_out/r2_run7_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_run7_pytest.log
tail -4 gpurun_out/r2_run7_pytest.log
for mode in scan scan4; do
GSR_BENCH_GRAPH=0 GSR_BWD_KERNEL=$mode python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_run7_bench_$mode.json 2> gpurun_out/r2_run7_bench_$mode.err; echo "bench $mode rc=$?"
done
python - <<'PY'
import json
for m in ("scan","scan4"):
    try:
        d=json.load(open(f"gpurun_out/r2_run7_bench_{m}.json")); print(m, d["value"], d["stages_ms"]["blend_bwd"])
    except Exception as e: print(m, "ERR", e)
PY
GSR_BENCH_GRAPH=0 GSR_BWD_KERNEL=scan4 ncu --set full --clock-control none --import-source on -k regex:blend_backward_scan -s 2 -c 1 -o gpurun_out/r2_run7_bwd_scan4 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2_run7_ncu1.log 2>&1
# one full-set pass over every own kernel of one view (second step)
GSR_BENCH_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:gsr -s 13 -c 13 -o gpurun_out/r2_run7_all python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_run7_ncu2.log 2>&1
ls -la gpurun_out/*.ncu-rep
