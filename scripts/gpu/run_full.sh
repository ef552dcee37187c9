cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -q -x -m gpu 2>&1 | tail -8
timeout 900 python bench.py > gpurun_out/r2_bench_full.json 2> gpurun_out/r2_bench_full.err; tail -3 gpurun_out/r2_bench_full.err; cat gpurun_out/r2_bench_full.json | cut -c1-1500
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8
