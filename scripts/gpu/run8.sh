cd $GRAFT_REPO_ROOT
export RDFC_UMMA_DBG=1
python scripts/prof_layer.py conv 32 64 128 228 304 3 2 0
python scripts/prof_layer.py conv 32 64 128 228 304 1 2 0
python scripts/prof_layer.py conv 32 128 256 114 152 3 2 0
python scripts/prof_layer.py conv 32 192 64 114 152 3 2 1
python scripts/prof_layer.py conv 32 64 64 228 304 3 1 0
