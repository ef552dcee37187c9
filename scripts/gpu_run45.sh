cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_nlspn.py -q -x -m gpu 2>&1 | tail -2
python scripts/prof_layer.py nlspn 32
python scripts/prof_layer.py nlspn 32
