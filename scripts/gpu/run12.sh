cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_esanet.py -q -x -m gpu 2>&1 | tail -8
timeout 300 python scripts/prof_c2_full.py 2>&1 | grep -v Warn | tee gpurun_out/r2_prof_c2_full2.txt | head -16
timeout 600 python bench.py --config c2 --no-cpu-baseline > gpurun_out/r2_c2_b.json 2> gpurun_out/r2_c2_b.err; tail -2 gpurun_out/r2_c2_b.err; cut -c1-300 gpurun_out/r2_c2_b.json
