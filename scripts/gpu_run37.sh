cd $GRAFT_REPO_ROOT
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench6.json 2> gpurun_out/bench6.err; python -c "
import json; d=json.load(open('gpurun_out/bench6.json')); print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','clocks','cpu_baseline')}); print(d['roofline']['frac'], d['roofline_dense']['achieved'])"; tail -3 gpurun_out/bench6.err
RDFC_NVCC_FLAGS=-DRDFC_UMMA_TIMERS python -m rdfc_gan_b200.build --force > /dev/null 2>&1
export RDFC_UMMA_DBG=1
for cfg in "32 64 64 228 304 3 1 0" "32 128 128 114 152 3 1 0" "32 128 160 228 304 3 1 0" "32 192 384 114 152 1 1 0"; do
  timeout 120 python scripts/prof_layer.py conv $cfg
done
timeout 300 python scripts/prof_plan.py 32 bf16 --timers "heads d.dec0" 2>&1 | grep -A12 "role timers"
