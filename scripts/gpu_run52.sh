cd $GRAFT_REPO_ROOT
L="32 128 160 228 304 3 1 0"
timeout 120 python scripts/prof_layer.py conv $L
RDFC_UMMA_BN=80 timeout 120 python scripts/prof_layer.py conv $L
RDFC_UMMA_NACC=1 timeout 120 python scripts/prof_layer.py conv $L
RDFC_UMMA_NACC=3 timeout 120 python scripts/prof_layer.py conv $L
L="32 128 96 228 304 3 1 0"
timeout 120 python scripts/prof_layer.py conv $L
RDFC_UMMA_NACC=1 timeout 120 python scripts/prof_layer.py conv $L
RDFC_UMMA_NACC=4 timeout 120 python scripts/prof_layer.py conv $L
L="32 64 64 228 304 3 1 0"
RDFC_UMMA_NACC=2 timeout 120 python scripts/prof_layer.py conv $L
RDFC_UMMA_SB=12 timeout 120 python scripts/prof_layer.py conv $L
