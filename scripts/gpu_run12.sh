set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_conv.py -q -x -m gpu 2>&1 | tail -3
L128="32 128 128 114 152 3 1 0"
L64="32 64 64 228 304 3 1 0"
python scripts/prof_layer.py conv $L128
RDFC_UMMA_SB=6 python scripts/prof_layer.py conv $L128
RDFC_UMMA_SB=12 python scripts/prof_layer.py conv $L128
RDFC_UMMA_NACC=4 python scripts/prof_layer.py conv $L128
RDFC_UMMA_NACC=4 RDFC_UMMA_SB=6 python scripts/prof_layer.py conv $L128
RDFC_UMMA_NACC=1 python scripts/prof_layer.py conv $L128
RDFC_UMMA_SA=2 python scripts/prof_layer.py conv $L128
python scripts/prof_layer.py conv $L64
RDFC_UMMA_SB=6 python scripts/prof_layer.py conv $L64
RDFC_UMMA_NACC=2 python scripts/prof_layer.py conv $L64
RDFC_UMMA_SA=2 python scripts/prof_layer.py conv $L64
python scripts/prof_layer.py conv 32 256 256 57 76 3 1 0
python scripts/prof_layer.py conv 32 512 512 29 38 3 1 0
python scripts/prof_layer.py conv 32 128 160 228 304 3 1 0
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 3 -c 1 -f -o gpurun_out/prof_umma_128b python scripts/prof_layer.py conv $L128 > gpurun_out/ncu_umma128.log 2>&1; tail -2 gpurun_out/ncu_umma128.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 3 -c 1 -f -o gpurun_out/prof_umma_64b python scripts/prof_layer.py conv $L64 > gpurun_out/ncu_umma64.log 2>&1; tail -2 gpurun_out/ncu_umma64.log
