cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_nlspn.py -q -x -m gpu 2>&1 | tail -2
python scripts/prof_layer.py nlspn 32
RDFC_NLSPN_PREFETCH=0 python scripts/prof_layer.py nlspn 32
RDFC_NLSPN_HALO=6 python scripts/prof_layer.py nlspn 32
RDFC_NLSPN_HALO=10 python scripts/prof_layer.py nlspn 32
RDFC_NLSPN_BAND_CTAS=456 python scripts/prof_layer.py nlspn 32
RDFC_NLSPN_BAND_CTAS=428 python scripts/prof_layer.py nlspn 32
