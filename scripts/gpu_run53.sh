cd $GRAFT_REPO_ROOT
python scripts/prof_layer.py nlspn 32
for mb in 32 56 80 120; do
  RDFC_NLSPN_GROUP_MB=$mb python scripts/prof_layer.py nlspn 32
  RDFC_NLSPN_EVICT_LAST=1 RDFC_NLSPN_GROUP_MB=$mb python scripts/prof_layer.py nlspn 32
done
