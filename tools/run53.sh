cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
R=r2_run53
L=gaussian-splatting-toolkit_b200/libgsr_b200.so
cp $L /tmp/lib_default.so
run() {  # name lib env...
  name=$1; lib=$2; shift 2
  if [ $lib = default ]; then cp /tmp/lib_default.so $L; else cp gpurun_variants/libgsr_$lib.so $L; fi
  env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --only-resident > gpurun_out/${R}_$name.json 2> gpurun_out/${R}_$name.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/${R}_$name.json").read().strip().splitlines()[-1])
s=d["stages_ms"]
print("%-12s %-24s step %.4f  bin %.4f  fwd %.4f  bwd %.4f" % ("$name", "$*", d["ms_per_step"], s["binning"], s["blend_fwd"], s["blend_bwd"]))
PY
}
run base default X=0
run na na X=0
run ef ef X=0
run trring2 trring2 X=0
cp /tmp/lib_default.so $L
(time timeout 600 python -m pytest tests/test_gpu_blend_adjoint_variants.py tests/test_gpu_parity.py -x -q -m gpu) > gpurun_out/${R}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${R}_pytest.log
