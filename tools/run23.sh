cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/profile_public_api.py async > gpurun_out/r2_run23_profile_async.txt 2>&1
python tools/profile_public_api.py sync > gpurun_out/r2_run23_profile_sync.txt 2>&1
python tools/host_overhead.py > gpurun_out/r2_run23_host_overhead.txt 2>&1
head -50 gpurun_out/r2_run23_profile_async.txt
