cd $GRAFT_REPO_ROOT
L64="32 64 64 228 304 3 1 0"
export RDFC_UMMA_DBG=1
RDFC_UMMA_SKIP=1 python scripts/prof_layer.py conv $L64 2>&1 | grep "epi\|conv B\|ACC_EMPTY"
RDFC_UMMA_SKIP=3 python scripts/prof_layer.py conv $L64 2>&1 | grep "epi\|conv B\|ACC_EMPTY"
unset RDFC_UMMA_DBG
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 3 -c 1 -f -o gpurun_out/prof_umma_64c python scripts/prof_layer.py conv $L64 > gpurun_out/ncu_umma64.log 2>&1; tail -2 gpurun_out/ncu_umma64.log
