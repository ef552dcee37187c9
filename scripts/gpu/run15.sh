cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_esanet.py tests/test_gpu_generator.py -q -x -m gpu 2>&1 | tail -4
timeout 600 python bench.py --config c2 --no-cpu-baseline --no-ref-gpu > gpurun_out/r2_c2_c.json 2> gpurun_out/r2_c2_c.err; tail -2 gpurun_out/r2_c2_c.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_c2_c.json").read().strip().splitlines()[-1])
print("c2 value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms", round(d["ms_per_step"],2))
PY
