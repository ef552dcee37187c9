cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_dcn.py -q -x -m gpu 2>&1 | tail -5
timeout 300 python scripts/prof_dcn.py 2>&1 | tee gpurun_out/r2_prof_dcn2.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_dcn_launches.csv python scripts/prof_dcn.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_dcn_launches.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
for r in rows[1:60]: print(r[ki][:60], r[vi])
PY
