cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_nlspn.py -q -x -m gpu 2>&1 | tail -3
for a in "32 228 304" "8 480 640" "256 228 304"; do timeout 120 python scripts/prof_nlspn.py $a 2>&1 | tail -1; done
