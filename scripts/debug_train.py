"""Development aid: per-parameter gradient errors of the training step against tests/golden/train_step.npz."""
import math, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
from make_train_golden import CASE, build_inputs, grad_sample_index, D_KW
from _synth import synth_state_dict
from rdfc_gan_b200.discriminator import PatchGANDiscriminator
from rdfc_gan_b200.generator import RDFGenerator
from rdfc_gan_b200.rdf_gan import RDFGAN
gold = np.load(os.path.join(ROOT, "tests/golden/train_step.npz"))
G = RDFGenerator(pretrained_on_imagenet=False, **CASE["kw"])
G.load_state_dict(synth_state_dict(G, seed=CASE["seed"], recipe="scaled", nlspn_stress=True))
D = PatchGANDiscriminator(**D_KW)
D.load_state_dict(synth_state_dict(D, seed=CASE["seed"] + 1, recipe="init"))   # init_weights(D), as rdf_gan.py:61
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
model = RDFGAN(G, D, device="cuda", args=CASE["args"])
model.train()
model.set_input(build_inputs())
model.forward()
if "--ref-fake" in sys.argv:        # feed the discriminator / losses the reference's own outputs: isolates D and the loss code
    model.fake_B_rgb_branch = torch.from_numpy(gold["out_depth_map_1"]).cuda().requires_grad_(True)
model.set_requires_grad(model.D, True); model.bucket_D.zero(); model.backward_D()
model.set_requires_grad(model.D, False); model.bucket_G.zero()
if "--ref-fake" not in sys.argv:
    model.backward_G()
if "--probe" in sys.argv:
    from make_train_golden import probe_tensors
    model.bucket_G.zero(); model.bucket_D.zero()
    model.forward()
    outs = (model.fake_B_rgb_branch, model.conf_map_rgb_branch, model.fake_B_depth_branch, model.conf_map_depth_branch, model.final_depth)
    sum((o * R.cuda()).sum() for o, R in zip(outs, probe_tensors())).backward()
rows = []
for net, mod in ((("P", model.G),) if "--probe" in sys.argv else (("G", model.G), ("D", model.D))):
    for name, p in mod.named_parameters():
        key = f"grad_{net}_{name}"
        if key not in gold.files or p.grad is None:
            continue
        idx = torch.from_numpy(grad_sample_index(name, p.numel()))
        got = p.grad.detach().reshape(-1).cpu()[idx].double(); want = torch.from_numpy(gold[key]).double()
        cos = float((got * want).sum() / (got.norm() * want.norm() + 1e-30))
        ratio = float(got.norm() / (want.norm() + 1e-30))
        rows.append((net, name, cos, ratio, float(want.norm())))
for r in rows:
    print(f"{r[0]} {r[1]:60s} cos {r[2]:+.4f}  |got|/|want| {r[3]:.4f}  |want| {r[4]:.3e}")
