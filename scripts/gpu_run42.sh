cd $GRAFT_REPO_ROOT
for shape in "32 512 512 29 38 3 1 0" "32 256 512 57 76 3 2 0" "32 512 512 29 38 3 2 0" "32 512 256 15 19 3 2 1" "32 768 128 29 38 3 2 1" "32 256 256 57 76 3 1 0" "32 128 256 114 152 3 2 0"; do
  for nacc in 1 2; do for bn in 128 256; do
    RDFC_UMMA_NACC=$nacc RDFC_UMMA_BN=$bn timeout 60 python scripts/prof_layer.py conv $shape 2>&1 | tail -1
  done; done
  timeout 60 python scripts/prof_layer.py conv $shape 2>&1 | tail -1
done
