cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_dcn.py -q -x -m gpu 2>&1 | tail -15
timeout 300 python scripts/prof_dcn.py 2>&1 | tee gpurun_out/r2_prof_dcn4.txt
