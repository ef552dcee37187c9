cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_generator.py -q -x -m gpu -k stream 2>&1 | tail -5
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench5.json 2> gpurun_out/bench5.err; python -c "
import json; d=json.load(open('gpurun_out/bench5.json')); print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','clocks')}); print(d['roofline']['frac'], d['roofline_dense']['achieved'])"; tail -3 gpurun_out/bench5.err
