cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
R=r2_run52
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
run trc100 default GSR_TR_CARVEOUT=100
run trc86 default GSR_TR_CARVEOUT=86
run trring1 trring1 X=0
run trring1c72 trring1 GSR_TR_CARVEOUT=72
run trring1c100 trring1 GSR_TR_CARVEOUT=100
run fwdc100 default GSR_FWD_CARVEOUT=100
run fwdc58 default GSR_FWD_CARVEOUT=58
run fwdring1 fwdring1 X=0
run fwdring1c44 fwdring1 GSR_FWD_CARVEOUT=44
run fwdring1c100 fwdring1 GSR_FWD_CARVEOUT=100
cp /tmp/lib_default.so $L
