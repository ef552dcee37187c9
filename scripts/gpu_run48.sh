cd $GRAFT_REPO_ROOT
timeout 120 python scripts/pair_check2.py
for cfg in "32 128 128 114 152 3 1 0" "32 64 64 228 304 3 1 0" "32 192 384 114 152 1 1 0" "32 192 64 114 152 3 2 1"; do
  timeout 120 python scripts/prof_layer.py conv $cfg
done
