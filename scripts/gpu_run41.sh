cd $GRAFT_REPO_ROOT
timeout 300 python scripts/prof_plan.py 32 bf16 --json gpurun_out/plan_steps_b32_v14.json > gpurun_out/plan_steps_b32_v14.log 2>&1; head -3 gpurun_out/plan_steps_b32_v14.log
