cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_generator.py -q -x -m gpu 2>&1 | tail -4
for cfg in "32 64 64 228 304 3 1 0" "32 128 128 114 152 3 1 0" "32 128 160 228 304 3 1 0" "32 192 384 114 152 1 1 0"; do
  timeout 120 python scripts/prof_layer.py conv $cfg
done
timeout 300 python scripts/prof_plan.py 32 bf16 --json gpurun_out/plan_steps_b32_v13.json 2>&1 | tee gpurun_out/plan_steps_b32_v13.log | head -3
