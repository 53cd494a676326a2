set -x
mkdir -p gpurun_out
python tools/cfg3_profile.py 60 > gpurun_out/cfg3_iter.log 2>&1; tail -1 gpurun_out/cfg3_iter.log
(timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ncu_cfg3.csv python tools/cfg3_profile.py 5) > gpurun_out/ncu_cfg3.log 2>&1; echo "ncu rc=$?"; tail -1 gpurun_out/ncu_cfg3.log
