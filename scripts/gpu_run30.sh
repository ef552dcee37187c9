cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_generator.py -q -x -m gpu 2>&1 | tail -4
timeout 300 python scripts/prof_plan.py 32 bf16 2>&1 | grep "stems\|plan B"
RDFC_NVCC_FLAGS=-DRDFC_UMMA_TIMERS python -m rdfc_gan_b200.build --force > /dev/null 2>&1
export RDFC_UMMA_DBG=1
for cfg in "32 64 64 228 304 3 1 0" "32 128 160 228 304 3 1 0" "32 256 256 57 76 3 1 0"; do
  timeout 120 python scripts/prof_layer.py conv $cfg
done
RDFC_UMMA_NACC=2 timeout 120 python scripts/prof_layer.py conv 32 64 64 228 304 3 1 0
