set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_generator.py -q -x -m gpu 2>&1 | tail -15
RDFC_UMMA_DBG=1 python scripts/prof_layer.py conv 8 64 64 228 304 3 1 0
RDFC_UMMA_NACC=2 python scripts/prof_layer.py conv 8 64 64 228 304 3 1 0
RDFC_UMMA_NACC=1 python scripts/prof_layer.py conv 8 64 64 228 304 3 1 0
RDFC_UMMA_DBG=1 python scripts/prof_layer.py conv 8 128 160 228 304 3 1 0
python scripts/prof_layer.py conv 8 128 128 114 152 3 1 0
python scripts/prof_layer.py conv 8 256 256 57 76 3 1 0
python scripts/prof_layer.py conv 8 512 512 29 38 3 1 0
python scripts/prof_layer.py conv 32 64 64 228 304 3 1 0
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench2.json 2> gpurun_out/bench2.err; cat gpurun_out/bench2.json; tail -5 gpurun_out/bench2.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 1 --warmup 3 --batch 8 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -3 gpurun_out/ncu_bench.log
