cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_nlspn.py -q -x -m gpu 2>&1 | tail -2
python scripts/prof_layer.py nlspn 32
RDFC_NLSPN_G=2 python scripts/prof_layer.py nlspn 32
RDFC_NLSPN_BAND=1 RDFC_NLSPN_PIX=2 python scripts/prof_layer.py nlspn 32
RDFC_NLSPN_BAND=1 RDFC_NLSPN_PIX=1 python scripts/prof_layer.py nlspn 32
RDFC_NLSPN_SIMPLE=1 python scripts/prof_layer.py nlspn 32
