set -x
cd $GRAFT_REPO_ROOT
RDFC_UMMA_DBG=1 python scripts/prof_layer.py conv 8 64 64 228 304 3 1 0
RDFC_UMMA_DBG=1 RDFC_UMMA_SB=8 python scripts/prof_layer.py conv 8 64 64 228 304 3 1 0
RDFC_UMMA_DBG=1 RDFC_UMMA_NACC=2 python scripts/prof_layer.py conv 8 64 64 228 304 3 1 0
RDFC_UMMA_DBG=1 python scripts/prof_layer.py conv 8 128 160 228 304 3 1 0
RDFC_UMMA_DBG=1 RDFC_UMMA_SB=8 python scripts/prof_layer.py conv 8 128 160 228 304 3 1 0
RDFC_UMMA_DBG=1 python scripts/prof_layer.py conv 32 512 512 29 38 3 1 0
RDFC_UMMA_DBG=1 python scripts/prof_layer.py conv 32 256 256 57 76 3 1 0
