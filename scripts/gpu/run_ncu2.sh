cd $GRAFT_REPO_ROOT
export CUDNN=0
timeout 600 ncu --set full --clock-control none -k regex:wgrad_umma_kernel -s 4 -c 3 -o /tmp/r2_wgrad -f python scripts/prof_wgrad.py > gpurun_out/r2_ncu_wgrad.log 2>&1
python scripts/ncu_summary.py /tmp/r2_wgrad.ncu-rep > gpurun_out/r2_ncu_wgrad_umma.md 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"dcn_|conv_umma|wgrad_umma_kernel" -s 0 -c 12 -o /tmp/r2_dcn -f python scripts/prof_dcn.py > gpurun_out/r2_ncu_dcn.log 2>&1
python scripts/ncu_summary.py /tmp/r2_dcn.ncu-rep > gpurun_out/r2_ncu_dcn.md 2>&1
grep -c "^##" gpurun_out/r2_ncu_wgrad_umma.md gpurun_out/r2_ncu_dcn.md
