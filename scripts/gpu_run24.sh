cd $GRAFT_REPO_ROOT
RDFC_NVCC_FLAGS=-DRDFC_UMMA_TIMERS python -m rdfc_gan_b200.build --force > /dev/null 2>&1
export RDFC_UMMA_DBG=1
for cfg in "32 64 64 228 304 3 1 0" "32 128 128 114 152 3 1 0" "32 192 384 114 152 1 1 0" "32 192 64 114 152 3 2 1" "32 512 512 29 38 3 1 0"; do
  timeout 120 python scripts/prof_layer.py conv $cfg
done
RDFC_UMMA_SKIP=1 timeout 120 python scripts/prof_layer.py conv 32 64 64 228 304 3 1 0
