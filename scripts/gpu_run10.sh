set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv
./scripts/umma_rate 2>&1 | tee gpurun_out/umma_rate2.log
python scripts/prof_plan.py 32 bf16 --json gpurun_out/plan_steps_b32.json 2>&1 | tee gpurun_out/plan_steps_b32.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_r1c.csv python scripts/prof_plan.py 8 bf16 > gpurun_out/ncu_plan.log 2>&1; tail -3 gpurun_out/ncu_plan.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 3 -c 1 -f -o gpurun_out/prof_umma_128 python scripts/prof_layer.py conv 32 128 128 114 152 3 1 0 > gpurun_out/ncu_umma128.log 2>&1; tail -2 gpurun_out/ncu_umma128.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 3 -c 1 -f -o gpurun_out/prof_umma_64 python scripts/prof_layer.py conv 32 64 64 228 304 3 1 0 > gpurun_out/ncu_umma64.log 2>&1; tail -2 gpurun_out/ncu_umma64.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nlspn_prop -s 60 -c 1 -f -o gpurun_out/prof_nlspn2 python scripts/prof_layer.py nlspn 32 > gpurun_out/ncu_nlspn2.log 2>&1; tail -2 gpurun_out/ncu_nlspn2.log
