set -x
cd $GRAFT_REPO_ROOT
RDFC_UMMA_DBG=1 python scripts/prof_layer.py conv 8 64 64 228 304 3 1 0
RDFC_UMMA_DBG=1 RDFC_UMMA_NACC=1 python scripts/prof_layer.py conv 8 64 64 228 304 3 1 0
RDFC_UMMA_DBG=1 python scripts/prof_layer.py conv 8 128 128 114 152 3 1 0
RDFC_UMMA_DBG=1 python scripts/prof_layer.py conv 8 512 512 29 38 3 1 0
python scripts/prof_layer.py conv 8 128 160 228 304 3 1 0
RDFC_UMMA_BN=128 python scripts/prof_layer.py conv 8 128 160 228 304 3 1 0
python scripts/prof_layer.py conv 8 256 256 57 76 3 1 0
RDFC_UMMA_BN=128 python scripts/prof_layer.py conv 8 256 256 57 76 3 1 0
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_generator.py -q -x -m gpu 2>&1 | tail -3
