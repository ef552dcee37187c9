cd $GRAFT_REPO_ROOT
timeout 300 python scripts/pair_check2.py
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_generator.py -q -x -m gpu 2>&1 | tail -4
for cfg in "32 64 64 228 304 3 1 0" "32 128 128 114 152 3 1 0" "32 128 160 228 304 3 1 0" "32 256 256 57 76 3 1 0" "32 512 512 29 38 3 1 0" "32 64 128 228 304 3 2 0" "32 192 64 114 152 3 2 1" "32 192 384 114 152 1 1 0"; do
  timeout 120 python scripts/prof_layer.py conv $cfg
  RDFC_UMMA_PAIR=0 timeout 120 python scripts/prof_layer.py conv $cfg
done
timeout 300 python scripts/prof_plan.py 32 bf16 2>&1 | head -2
