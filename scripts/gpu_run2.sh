set -x
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -q --maxfail=10 -m gpu > gpurun_out/t_all.log 2>&1; tail -40 gpurun_out/t_all.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -12 gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err; cat gpurun_out/bench1.json; tail -5 gpurun_out/bench1.err
timeout 600 python bench.py --steps 5 --warmup 3 --precision fp32 --batch 4 --no-cpu-baseline > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err; cat gpurun_out/bench_fp32.json; tail -5 gpurun_out/bench_fp32.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 3 --batch 8 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -3 gpurun_out/ncu_bench.log; wc -l gpurun_out/launches_r1.csv
