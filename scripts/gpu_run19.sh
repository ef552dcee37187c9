set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_nlspn.py tests/test_gpu_generator.py -q -x -m gpu 2>&1 | tail -5
python scripts/prof_layer.py nlspn 32
RDFC_NLSPN_SIMPLE=1 python scripts/prof_layer.py nlspn 32
RDFC_NLSPN_HALO=6 python scripts/prof_layer.py nlspn 32
RDFC_NLSPN_HALO=12 python scripts/prof_layer.py nlspn 32
RDFC_NLSPN_BAND_CTAS=296 python scripts/prof_layer.py nlspn 32
RDFC_NLSPN_BAND_CTAS=592 python scripts/prof_layer.py nlspn 32
RDFC_NLSPN_BAND_CTAS=888 python scripts/prof_layer.py nlspn 32
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nlspn_prop -s 60 -c 1 -f -o gpurun_out/prof_nlspn_band python scripts/prof_layer.py nlspn 32 > gpurun_out/ncu_nlspn_band.log 2>&1; tail -2 gpurun_out/ncu_nlspn_band.log
