# ncu evidence for the blend kernels (run under gpurun, 1 GPU). Numbers printed under ncu are never bench values.
set -x
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_bench.log 2>&1
echo "launch list rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:blend_ -s 2 -c 2 -f -o gpurun_out/prof_blend \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_bench.log 2>&1
echo "full rc=$?"
ls -la gpurun_out | tail
