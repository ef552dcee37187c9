cd $GRAFT_REPO_ROOT
N=${1:-8}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; }
timeout 600 bash -c "$(declare -f run); N=$N; run --config c3 --no-cpu-baseline --no-ref-gpu" > gpurun_out/r2_c3_n$N.json 2> gpurun_out/r2_c3_n$N.err; tail -2 gpurun_out/r2_c3_n$N.err; cut -c1-600 gpurun_out/r2_c3_n$N.json
timeout 600 bash -c "$(declare -f run); N=$N; run --config c4 --steps 8 --warmup 3 --no-cpu-baseline" > gpurun_out/r2_c4_n$N.json 2> gpurun_out/r2_c4_n$N.err; tail -2 gpurun_out/r2_c4_n$N.err; cut -c1-400 gpurun_out/r2_c4_n$N.json; grep -o '"allreduce": {[^}]*}' gpurun_out/r2_c4_n$N.json
timeout 600 bash -c "$(declare -f run); N=$N; run --config c5 --no-cpu-baseline --no-ref-gpu" > gpurun_out/r2_c5_n$N.json 2> gpurun_out/r2_c5_n$N.err; tail -2 gpurun_out/r2_c5_n$N.err; cut -c1-600 gpurun_out/r2_c5_n$N.json
