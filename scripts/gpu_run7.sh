set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_nlspn.py -q -x -m gpu 2>&1 | tail -5
python scripts/prof_layer.py nlspn 32
RDFC_NLSPN_CTAS_PER_SM=3 python scripts/prof_layer.py nlspn 32
RDFC_NLSPN_STAGES=2 RDFC_NLSPN_CTAS_PER_SM=3 python scripts/prof_layer.py nlspn 32
RDFC_NLSPN_STAGES=2 RDFC_NLSPN_CTAS_PER_SM=2 python scripts/prof_layer.py nlspn 32
timeout 300 python scripts/check_layers.py 32 2>&1 | grep -v "mbarrier timeout" | grep -E "FAILED|step" | awk '{print $2, $4}' | tr '\n' ' '
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench2.json 2> gpurun_out/bench2.err; cat gpurun_out/bench2.json; tail -5 gpurun_out/bench2.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:rdfc -c 400 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 1 --warmup 3 --batch 32 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -3 gpurun_out/ncu_bench.log | cut -c1-300
