# A/B of one RDFC_* knob on the plan's per-step times:  bash scripts/gpu/run_ab.sh KNOB "pattern|pattern"
cd $GRAFT_REPO_ROOT
K=$1; PAT=$2
for v in 1 0; do echo "== $K=$v"; env $K=$v timeout 300 python scripts/prof_plan.py 32 2>&1 | grep -E "plan B|$PAT"; done
