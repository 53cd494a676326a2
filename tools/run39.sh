set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
R=r2_run39
L=gaussian-splatting-toolkit_b200/libgsr_b200.so
cp $L /tmp/lib_default.so
for v in default u8_c4 u4_c5 u2_c4; do
  if [ $v = default ]; then cp /tmp/lib_default.so $L; else cp gpurun_variants_libgsr_$v.so $L; fi
  python bench.py --steps 30 --warmup 5 --only-resident > gpurun_out/${R}_bench_$v.json 2> gpurun_out/${R}_bench_$v.err; echo "bench $v rc=$?"; cat gpurun_out/${R}_bench_$v.json | cut -c1-420
done
cp /tmp/lib_default.so $L
