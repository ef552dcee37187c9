cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_train.py -q -x -m gpu 2>&1 | tail -3
timeout 600 python scripts/prof_train.py > gpurun_out/r2_prof_train_c.txt 2>&1; head -50 gpurun_out/r2_prof_train_c.txt
