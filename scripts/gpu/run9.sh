cd $GRAFT_REPO_ROOT
export CUDNN=0
python scripts/prof_wgrad.py 2>&1 | cut -c1-75
for dbg in 1 2 3; do echo "== RDFC_WGRAD_DBG=$dbg"; RDFC_WGRAD_DBG=$dbg NLAYERS=5 python scripts/prof_wgrad.py 2>&1 | cut -c1-75; done
echo "== TW=32"; RDFC_WGRAD_TW=32 NLAYERS=5 python scripts/prof_wgrad.py 2>&1 | cut -c1-75
timeout 600 python -m pytest tests/test_gpu_train.py -q -x -m gpu -k "filter_gradient" 2>&1 | tail -3
