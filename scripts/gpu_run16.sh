cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv
L64="32 64 64 228 304 3 1 0"
export RDFC_UMMA_DBG=1
python scripts/prof_layer.py conv $L64 2>&1 | grep "epi\|conv B\|ACC_EMPTY"
RDFC_UMMA_SKIP=1 python scripts/prof_layer.py conv $L64 2>&1 | grep "epi\|conv B\|ACC_EMPTY"
RDFC_UMMA_SKIP=2 python scripts/prof_layer.py conv $L64 2>&1 | grep "epi\|conv B\|ACC_EMPTY"
RDFC_UMMA_SKIP=3 python scripts/prof_layer.py conv $L64 2>&1 | grep "epi\|conv B\|ACC_EMPTY"
