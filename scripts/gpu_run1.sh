set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
cd $GRAFT_REPO_ROOT
python -c "import torch; print(torch.cuda.get_device_name(0))"
timeout 900 python -m pytest tests/test_gpu_dcn.py -q --maxfail=20 -x -m gpu > gpurun_out/t_dcn.log 2>&1; tail -30 gpurun_out/t_dcn.log
timeout 600 python -m pytest tests/test_gpu_nlspn.py -q --maxfail=20 -m gpu > gpurun_out/t_nlspn.log 2>&1; tail -30 gpurun_out/t_nlspn.log
timeout 600 python -m pytest tests/test_gpu_conv.py -q --maxfail=20 -m gpu -k "False or small or instnorm" > gpurun_out/t_conv32.log 2>&1; tail -30 gpurun_out/t_conv32.log
timeout 600 python -m pytest tests/test_gpu_conv.py -q --maxfail=20 -m gpu -k "True" > gpurun_out/t_conv16.log 2>&1; tail -40 gpurun_out/t_conv16.log
timeout 900 python -m pytest tests/test_gpu_generator.py -q --maxfail=20 -m gpu -k "fp32 or oracle or dcvgan" > gpurun_out/t_gen32.log 2>&1; tail -40 gpurun_out/t_gen32.log
