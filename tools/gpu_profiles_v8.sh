# round-1 final profile evidence for the current kernels: launch list of the bench command, --set full of the two blend
# kernels, launch list of a cfg3-style training iteration
set -x
mkdir -p gpurun_out
(timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ncu_launches_v8.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline) > gpurun_out/ncu_launches_v8.log 2>&1; echo "launches rc=$?"
(timeout 500 ncu --set full --clock-control none --import-source on -k 'regex:^blend_(forward|backward)_kernel$' -s 2 -c 2 -f -o gpurun_out/prof_blend_v8 python bench.py --steps 2 --warmup 1 --no-cpu-baseline) > gpurun_out/ncu_blend_v8.log 2>&1; echo "full rc=$?"
python tools/cfg3_profile.py 60 > gpurun_out/cfg3_iter.log 2>&1; tail -1 gpurun_out/cfg3_iter.log
(timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ncu_cfg3.csv python tools/cfg3_profile.py 5) > gpurun_out/ncu_cfg3.log 2>&1; echo "cfg3 ncu rc=$?"
ls -la gpurun_out | tail -8
