cd $GRAFT_REPO_ROOT
timeout 120 python scripts/pair_check2.py | tail -2
RDFC_NVCC_FLAGS=-DRDFC_UMMA_TIMERS python -m rdfc_gan_b200.build --force > /dev/null 2>&1
L="32 128 128 114 152 3 1 0"
RDFC_UMMA_DBG=1 timeout 120 python scripts/prof_layer.py conv $L
RDFC_UMMA_SB=8 timeout 120 python scripts/prof_layer.py conv $L | tail -1
RDFC_UMMA_SB=12 timeout 120 python scripts/prof_layer.py conv $L | tail -1
RDFC_UMMA_PAIR=0 RDFC_UMMA_DBG=1 timeout 120 python scripts/prof_layer.py conv $L
