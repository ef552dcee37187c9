cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_conv.py -q -x -m gpu -k "stems or wadain_conv" 2>&1 | tail -15
