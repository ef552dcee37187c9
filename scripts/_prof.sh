timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_bench.csv \
  python bench.py --steps 2 --warmup 3 --batch 32 --no-cpu-baseline --no-ref-gpu --no-parity > gpurun_out/r2_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nlspn_prop_packed -s 20 -c 2 -o gpurun_out/r2_nlspn_packed -f \
  python scripts/prof_nlspn.py 32 > gpurun_out/r2_ncu_nlspn.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nlspn_prop_band -s 20 -c 1 -o gpurun_out/r2_nlspn_band -f \
  python scripts/prof_nlspn.py 32 >> gpurun_out/r2_ncu_nlspn.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:conv_umma -s 200 -c 40 -o /tmp/r2_conv_umma -f \
  python scripts/prof_plan.py 32 > gpurun_out/r2_ncu_conv.log 2>&1
python scripts/ncu_summary.py /tmp/r2_conv_umma.ncu-rep > gpurun_out/r2_ncu_conv_umma.md 2>&1
python scripts/ncu_summary.py gpurun_out/r2_nlspn_packed.ncu-rep gpurun_out/r2_nlspn_band.ncu-rep > gpurun_out/r2_ncu_nlspn.md 2>&1
du -sh gpurun_out; ls gpurun_out | head -30
