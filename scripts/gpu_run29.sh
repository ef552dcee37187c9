cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_generator.py -q -x -m gpu 2>&1 | tail -8
for cfg in "32 64 64 228 304 3 1 0" "32 128 128 114 152 3 1 0" "32 256 256 57 76 3 1 0" "32 512 512 29 38 3 1 0" "32 512 512 29 38 3 2 0" "32 512 256 15 19 3 2 1"; do
  timeout 120 python scripts/prof_layer.py conv $cfg
done
RDFC_UMMA_NAX=1 timeout 120 python scripts/prof_layer.py conv 32 64 64 228 304 3 1 0
RDFC_UMMA_NAX=2 timeout 120 python scripts/prof_layer.py conv 32 64 64 228 304 3 1 0
timeout 300 python scripts/prof_plan.py 32 bf16 --json gpurun_out/plan_steps_b32_v10.json 2>&1 | tee gpurun_out/plan_steps_b32_v10.log | head -3
