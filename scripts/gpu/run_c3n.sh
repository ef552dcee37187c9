cd $GRAFT_REPO_ROOT
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --config c3 --no-cpu-baseline --no-ref-gpu > gpurun_out/r2_c3_n$N.json 2> gpurun_out/r2_c3_n$N.err; tail -2 gpurun_out/r2_c3_n$N.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_c3_n$N.json").read().strip().splitlines()[-1])
print("N=$N value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms", round(d["ms_per_step"],2), d["clocks"])
PY
