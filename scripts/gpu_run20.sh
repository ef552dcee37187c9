cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_nlspn.py -q -x -m gpu 2>&1 | tail -2
python scripts/prof_layer.py nlspn 32
RDFC_NLSPN_PIX=2 python scripts/prof_layer.py nlspn 32
RDFC_NLSPN_HALO=6 python scripts/prof_layer.py nlspn 32
RDFC_NLSPN_BAND_CTAS=592 python scripts/prof_layer.py nlspn 32
RDFC_NLSPN_BAND_CTAS=1480 python scripts/prof_layer.py nlspn 32
