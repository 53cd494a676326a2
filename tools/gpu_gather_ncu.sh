mkdir -p gpurun_out
timeout 120 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:densify_gather -c 2 --csv --log-file gpurun_out/ncu_gather_v3.csv python tools/f2_profile.py > gpurun_out/ncu_gather_v3.log 2>&1; echo "rc=$?"; grep -c gather gpurun_out/ncu_gather_v3.csv
