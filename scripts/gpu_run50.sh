cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_generator.py -q -x -m gpu 2>&1 | tail -4
timeout 300 python scripts/prof_plan.py 32 bf16 2>&1 | head -2
