cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_nlspn.py -q -x -m gpu 2>&1 | tail -3
for k in 0 1; do echo "== RDFC_NLSPN_PITCH32=$k"; for a in "32 228 304" "8 480 640" "256 228 304"; do RDFC_NLSPN_PITCH32=$k timeout 120 python scripts/prof_nlspn.py $a 2>&1 | tail -2; done; done
echo "== smooth offsets (OFFSCALE small)"; for k in 0 1; do RDFC_NLSPN_PITCH32=$k OFFSCALE=0.05 timeout 120 python scripts/prof_nlspn.py 32 228 304 2>&1 | tail -1; done
