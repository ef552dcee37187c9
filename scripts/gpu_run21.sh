cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_nlspn.py tests/test_gpu_generator.py -q -x -m gpu 2>&1 | tail -3
python scripts/prof_layer.py nlspn 32
RDFC_NLSPN_BAND=1 RDFC_NLSPN_PIX=2 python scripts/prof_layer.py nlspn 32
RDFC_NLSPN_STAGES=2 python scripts/prof_layer.py nlspn 32
RDFC_NLSPN_G=1 python scripts/prof_layer.py nlspn 32
RDFC_NLSPN_G=1 RDFC_NLSPN_STAGES=4 python scripts/prof_layer.py nlspn 32
RDFC_NLSPN_HALO=6 python scripts/prof_layer.py nlspn 32
RDFC_NLSPN_RING_CTAS=296 python scripts/prof_layer.py nlspn 32
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nlspn_prop -s 60 -c 1 -f -o gpurun_out/prof_nlspn_ring python scripts/prof_layer.py nlspn 32 > gpurun_out/ncu_nlspn_ring.log 2>&1; tail -2 gpurun_out/ncu_nlspn_ring.log
