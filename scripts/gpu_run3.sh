set -x
cd $GRAFT_REPO_ROOT
./scripts/umma_rate 2>&1 | tee gpurun_out/umma_rate.log
python scripts/prof_layer.py conv 8 64 64 228 304 3 1 0
RDFC_UMMA_SB=8 python scripts/prof_layer.py conv 8 64 64 228 304 3 1 0
RDFC_UMMA_NACC=2 python scripts/prof_layer.py conv 8 64 64 228 304 3 1 0
RDFC_UMMA_NACC=1 python scripts/prof_layer.py conv 8 64 64 228 304 3 1 0
python scripts/prof_layer.py conv 8 128 128 114 152 3 1 0
python scripts/prof_layer.py conv 8 512 512 29 38 3 1 0
python scripts/prof_layer.py conv 32 64 64 228 304 3 1 0
python scripts/prof_layer.py nlspn 32
RDFC_NLSPN_GROUP_MB=32 python scripts/prof_layer.py nlspn 32
RDFC_NLSPN_GROUP_MB=100000 python scripts/prof_layer.py nlspn 32
python scripts/prof_layer.py nlspn 8
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 3 -c 1 -f -o gpurun_out/prof_umma_en2 python scripts/prof_layer.py conv 8 64 64 228 304 3 1 0 > gpurun_out/ncu_umma.log 2>&1; tail -2 gpurun_out/ncu_umma.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nlspn_prop -s 60 -c 1 -f -o gpurun_out/prof_nlspn python scripts/prof_layer.py nlspn 32 > gpurun_out/ncu_nlspn.log 2>&1; tail -2 gpurun_out/ncu_nlspn.log
ls -la gpurun_out/
