cd $GRAFT_REPO_ROOT
timeout 300 python scripts/prof_layer.py conv 32 128 128 114 152 3 1 0 2>&1 | tail -3
RDFC_UMMA_PAIR=0 timeout 300 python scripts/prof_layer.py conv 32 128 128 114 152 3 1 0 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_conv.py -q -x -m gpu 2>&1 | tail -8
