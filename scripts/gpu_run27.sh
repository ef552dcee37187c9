cd $GRAFT_REPO_ROOT
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench4.json 2> gpurun_out/bench4.err; cut -c1-1800 gpurun_out/bench4.json; tail -3 gpurun_out/bench4.err
