set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
R=r2_run32
python -m pytest tests/test_gpu_binning_device.py -m gpu -q -x > gpurun_out/${R}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${R}_pytest.log
tail -25 gpurun_out/${R}_pytest.log | cut -c1-300
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/${R}_bench.err | cut -c1-400; python -c "
import json;d=json.loads([l for l in open('gpurun_out/${R}_bench.json') if l.startswith('{')][-1]);print(d['value'], d['e2e']['value'], d['e2e']['mode'], d['e2e']['eager'], d['cuda_graph']['value'])"
