cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_train.py -q -x -m gpu -k "filter_gradient" 2>&1 | tail -15
timeout 600 python scripts/prof_wgrad.py 2>&1 | tee gpurun_out/r2_prof_wgrad.txt
