cd $GRAFT_REPO_ROOT
export CUDNN=0
timeout 600 python -m pytest tests/test_gpu_train.py -q -x -m gpu -k "filter_gradient" 2>&1 | tail -3
python scripts/prof_wgrad.py 2>&1 | cut -c1-150
echo "== RDFC_WGRAD_DBG=1"; RDFC_WGRAD_DBG=1 NLAYERS=5 python scripts/prof_wgrad.py 2>&1 | cut -c1-75
