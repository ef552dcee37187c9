timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6
timeout 400 python bench.py --config c2 --steps 10 2>gpurun_out/r2_c2.err | tail -1 > gpurun_out/r2_c2.json
python - <<'PY'
import json
try:
    j=json.load(open("gpurun_out/r2_c2.json")); print({k: j.get(k) for k in ("value","ms_per_step","e2e","value_fp32","value_fp32_tc","ref_gpu","cpu_baseline","gpu_launches")}); print(j["parity"])
except Exception as e: print("c2 failed", e)
PY
tail -3 gpurun_out/r2_c2.err
timeout 400 python bench.py --config c5 --steps 5 --no-cpu-baseline 2>gpurun_out/r2_c5.err | tail -1 > gpurun_out/r2_c5.json
python - <<'PY'
import json
try:
    j=json.load(open("gpurun_out/r2_c5.json")); print({k: j.get(k) for k in ("value","ms_per_step","e2e","value_fp32_tc","ref_gpu")}); print(j["parity"]["pred_depth"], j["roofline"]["frac"])
except Exception as e: print("c5 failed", e)
PY
tail -3 gpurun_out/r2_c5.err
