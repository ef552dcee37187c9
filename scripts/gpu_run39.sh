cd $GRAFT_REPO_ROOT
python - <<'PY'
import sys, torch, time
sys.path.insert(0, 'tests')
import bench
from _synth import synth_inputs, synth_state_dict
from rdfc_gan_b200.generator import RDFGenerator
from oracle import generator as ogen
nl = dict(prop_kernel=3, prop_time=12, affinity="TGASS", affinity_gamma=0.5, conf_prop=True, preserve_input=False)
G = RDFGenerator(pretrained_on_imagenet=False, use_nlspn_refine=True, nlspn_configs=nl).eval()
G.load_state_dict(synth_state_dict(G, seed=3, recipe="scaled", nlspn_stress=True))
rgb, normal, depth = synth_inputs(1, 480, 640, seed=3)
ref = ogen.generator_forward(G.state_dict(), normal, depth, use_nlspn_refine=True, nlspn_configs=nl)
G = G.cuda()
with torch.no_grad():
    o32 = G.set_precision("fp32")(rgb.cuda(), depth.cuda(), normal.cuda())
    o16 = G.set_precision("bf16")(rgb.cuda(), depth.cuda(), normal.cuda())
for k, r in ref.items():
    print(f"480x640 {k}: fp32 max-abs {float((o32[k].cpu()-r).abs().max()):.2e}  bf16 max-abs {float((o16[k].cpu()-r).abs().max()):.2e} rmse {float(((o16[k].cpu()-r)**2).mean().sqrt()):.2e}")
# throughput at 480x640, B=8, bf16
rgb, normal, depth = synth_inputs(8, 480, 640, seed=4)
r, n, d = rgb.cuda(), normal.cuda(), depth.cuda()
with torch.no_grad():
    for _ in range(3): G(r, d, n)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): G(r, d, n)
    torch.cuda.synchronize()
print(f"480x640 B=8 bf16 12 it.: {8*10/(time.perf_counter()-t0):.1f} maps/s")
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 2>&1 | tail -2 | cut -c1-700
