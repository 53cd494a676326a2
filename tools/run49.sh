cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
R=r2_run49
run() {  # name workload steps env...
  name=$1; wl=$2; st=$3; shift 3
  env "$@" timeout 400 python bench.py --steps $st --warmup 5 --only-resident --workload $wl > gpurun_out/${R}_$name.json 2> gpurun_out/${R}_$name.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/${R}_$name.json").read().strip().splitlines()[-1])
print("$name", "$*", round(d["ms_per_step"],4), "binning", round(d["stages_ms"]["binning"],4))
PY
}
run A cfg2 30 GSR_BIN_ROUND=32 GSR_BIN_EARLY=0 GSR_FILL_EARLY=1
run B cfg2 30 GSR_BIN_ROUND=1024 GSR_BIN_EARLY=0 GSR_FILL_EARLY=1
run C cfg2 30 GSR_BIN_ROUND=32 GSR_BIN_EARLY=1 GSR_FILL_EARLY=1
run D cfg2 30 GSR_BIN_ROUND=32 GSR_BIN_EARLY=0 GSR_FILL_EARLY=0
run E cfg2 30 GSR_BIN_ROUND=1024 GSR_BIN_EARLY=0 GSR_FILL_EARLY=0
run F cfg2 30 GSR_BIN_ROUND=1024 GSR_BIN_EARLY=1 GSR_FILL_EARLY=1
run A4 cfg4 10 GSR_BIN_ROUND=32 GSR_BIN_EARLY=0 GSR_FILL_EARLY=1
run B4 cfg4 10 GSR_BIN_ROUND=1024 GSR_BIN_EARLY=0 GSR_FILL_EARLY=1
run E4 cfg4 10 GSR_BIN_ROUND=1024 GSR_BIN_EARLY=0 GSR_FILL_EARLY=0
