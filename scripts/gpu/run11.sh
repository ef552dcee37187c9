cd $GRAFT_REPO_ROOT
for cfg in "32 128 160 228 304 3 1 0" "32 128 96 228 304 3 1 0" "32 256 256 57 76 3 1 0"; do
  for nacc in 0 1 2 3; do
    if [ $nacc = 0 ]; then python scripts/prof_layer.py conv $cfg 2>&1 | tail -1; else RDFC_UMMA_NACC=$nacc python scripts/prof_layer.py conv $cfg 2>&1 | tail -1; fi
  done
done
