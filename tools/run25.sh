set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
R=r2_run25
python -m pytest tests/test_gpu_parity.py tests/test_gpu_blend_adjoint_variants.py tests/test_gpu_fused.py tests/test_gpu_properties.py -m gpu -q -x > gpurun_out/${R}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${R}_pytest.log
tail -3 gpurun_out/${R}_pytest.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err; echo "bench rc=$?"; python -c "
import json;d=json.loads([l for l in open('gpurun_out/${R}_bench.json') if l.startswith('{')][-1]);print(d['value'], d['stages_ms'], d['fused_operator']['value'], d['e2e']['value'], d['cuda_graph']['value'])"
