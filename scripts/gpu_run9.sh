set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_conv.py -q -x -m gpu 2>&1 | tail -5
timeout 600 python -m pytest tests/test_gpu_generator.py -q -x -m gpu 2>&1 | tail -5
python scripts/prof_layer.py conv 8 64 64 228 304 3 1 0
python scripts/prof_layer.py conv 32 64 64 228 304 3 1 0
RDFC_UMMA_NSETS=1 python scripts/prof_layer.py conv 32 64 64 228 304 3 1 0
python scripts/prof_layer.py conv 32 128 160 228 304 3 1 0
python scripts/prof_layer.py conv 32 128 96 228 304 3 1 0
python scripts/prof_layer.py conv 32 128 128 114 152 3 1 0
python scripts/prof_layer.py conv 32 256 256 57 76 3 1 0
RDFC_UMMA_BN=256 python scripts/prof_layer.py conv 32 256 256 57 76 3 1 0
python scripts/prof_layer.py conv 32 512 512 29 38 3 1 0
python scripts/prof_layer.py conv 32 64 128 228 304 3 2 0
python scripts/prof_layer.py conv 32 192 64 114 152 3 2 1
python scripts/prof_layer.py nlspn 32
RDFC_NLSPN_ROWS=1 python scripts/prof_layer.py nlspn 32
RDFC_NLSPN_ROWS=1 RDFC_NLSPN_CTAS_PER_SM=2 python scripts/prof_layer.py nlspn 32
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench3.json 2> gpurun_out/bench3.err; cat gpurun_out/bench3.json | cut -c1-1500; tail -5 gpurun_out/bench3.err
