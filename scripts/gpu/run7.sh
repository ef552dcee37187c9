cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_conv.py -q -x -m gpu 2>&1 | tail -6
RDFC_UMMA_DENSE2=0 timeout 300 python scripts/prof_plan.py 32 > gpurun_out/r2_plan_dense0.txt 2>&1
RDFC_UMMA_DENSE2=1 timeout 300 python scripts/prof_plan.py 32 > gpurun_out/r2_plan_dense1.txt 2>&1
head -1 gpurun_out/r2_plan_dense0.txt; grep " s2 T0" gpurun_out/r2_plan_dense0.txt
head -1 gpurun_out/r2_plan_dense1.txt; grep " s2 T0" gpurun_out/r2_plan_dense1.txt
