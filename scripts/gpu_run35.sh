cd $GRAFT_REPO_ROOT
RDFC_NVCC_FLAGS=-DRDFC_UMMA_TIMERS python -m rdfc_gan_b200.build --force > /dev/null 2>&1
export RDFC_UMMA_DBG=1
timeout 300 python scripts/prof_plan.py 32 bf16 --timers "heads d.dec0" 2>&1 | grep -A12 "role timers"
timeout 300 python scripts/prof_plan.py 32 bf16 --timers "stems" 2>&1 | grep -A12 "role timers"
timeout 300 python scripts/prof_plan.py 32 bf16 --timers "wadain_conv fuse4" 2>&1 | grep -A12 "role timers"
timeout 300 python scripts/prof_plan.py 32 bf16 --timers "r.de2" 2>&1 | grep -A12 "role timers"
