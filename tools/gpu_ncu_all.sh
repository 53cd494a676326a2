# final ncu evidence of the round: --set full captures of every own kernel of one view (the launch list comes from
# tools/gpu_ncu_blend.sh)
set -x
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on \
    -k 'regex:^(bin_edges_u32|blend_backward|blend_forward|count_tiles|depth_keys|emit_sorted|project_backward|project_forward|sh_backward|sh_forward)_kernel$' \
    -s 10 -c 10 -f -o gpurun_out/prof_all \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_bench.log 2>&1
echo "full rc=$?"
ls -la gpurun_out | tail -5
