cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_train.py -q -x -m gpu 2>&1 | tail -8
timeout 600 python bench.py --config c4 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2_c4_b.json 2> gpurun_out/r2_c4_b.err; tail -3 gpurun_out/r2_c4_b.err; cat gpurun_out/r2_c4_b.json
timeout 600 python scripts/prof_train.py 2>&1 | tee gpurun_out/r2_prof_train_b.txt | head -40
