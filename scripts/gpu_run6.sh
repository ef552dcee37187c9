set -x
cd $GRAFT_REPO_ROOT
timeout 300 python scripts/check_layers.py 32 2>&1 | grep -v "mbarrier timeout" | tail -150
timeout 600 python -m pytest tests/test_gpu_nlspn.py -q -x -m gpu 2>&1 | tail -5
python scripts/prof_layer.py nlspn 32
RDFC_NLSPN_CTAS_PER_SM=1 python scripts/prof_layer.py nlspn 32
RDFC_NLSPN_CTAS_PER_SM=3 python scripts/prof_layer.py nlspn 32
RDFC_NLSPN_SIMPLE=1 python scripts/prof_layer.py nlspn 32
python scripts/prof_layer.py nlspn 8
