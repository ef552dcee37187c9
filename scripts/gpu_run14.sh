set -x
cd $GRAFT_REPO_ROOT
export RDFC_UMMA_DBG=1
for cfg in "32 64 64 228 304 3 1 0" "32 128 128 114 152 3 1 0" "32 512 512 29 38 3 1 0" "32 192 384 114 152 1 1 0" "32 192 64 114 152 3 2 1" "32 128 160 228 304 3 1 0"; do
  timeout 120 python scripts/prof_layer.py conv $cfg
done
