#!/usr/bin/env python
"""bench.py -- depth maps/s of the generator forward on N B200s, with the NLSPN roofline and a CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch 32] [--precision bf16|fp32] [--impl reference]

A step = one generator forward (RDFC-GAN RDFGenerator: 2 x ResNet-18 encoder/decoder, W-AdaIN fusion, NLSPN TGASS
18 iterations) over one synthetic batch at 228x304.  N > 1: launched by torchrun, one rank per GPU, each rank runs the
same per-GPU batch (weak scaling, the path shards by image: no collective on the data path); the timed region is
bracketed by a barrier + synchronize and the max over ranks is taken.  Rank 0 prints ONE JSON line.

  value    : maps/s with inputs resident in HBM (CUDA-graph replay of the plan), CUDA-event timed per step,
             L2 flushed between steps.
  e2e      : the same metric through the public module call G(rgb, depth, normal) with pinned HOST inputs: the
             H2D copies and the D2H read of pred_depth are inside the timed region.
  roofline : the NLSPN propagation kernel timed alone (CUDA events on the launch stream):
             achieved = 116 B x pixels per launch / mean launch time, against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline / --impl reference : the oracle port (oracle/generator.py: torch CPU + the C DCN oracle) on all host
             cores, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

H, W = 228, 304
NLSPN_CFG = dict(prop_kernel=3, prop_time=18, affinity="TGASS", affinity_gamma=0.5, conf_prop=True, preserve_input=False)
ALGO_BYTES_PER_PIXEL_ITER = 116          # 18 offsets + 9 affinities + 1 feature read, 1 feature written, fp32
GFLOP_PER_IMAGE = 221.4                  # 2 x MACs of every Conv/ConvT/Linear at 228x304 (ResNet-18 config)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], bf16=p["bf16_tflops"], bf16_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    src="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, src="fallback")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md recipe): one streaming
    `nvidia-smi -lms 50` process, so that even a 0.2 s timed region yields several samples."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if line.strip():
                    self.rows.append([c.strip() for c in line.strip().split(",")])
        except Exception:
            pass

    def summary(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=6)
        sm = [int(r[0]) for r in self.rows if r[0].isdigit()]
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return dict(sm_mhz=int(statistics.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons,
                    samples=len(sm))


def build_generator():
    import torch
    from _synth import synth_state_dict
    from rdfc_gan_b200.generator import RDFGenerator
    G = RDFGenerator(pretrained_on_imagenet=False, use_nlspn_refine=True, nlspn_configs=NLSPN_CFG).eval()
    # the reference's random init (init_weights) + NLSPN offsets of trained magnitude (sigma ~ 2 px), SURVEY 8d
    G.load_state_dict(synth_state_dict(G, seed=0, recipe="init", nlspn_stress=True))
    return G


def cpu_baseline(n_images, reps):
    """Oracle port on the host cores: returns (maps/s, cores, seconds per forward)."""
    import torch
    from _synth import synth_inputs
    from oracle import generator as ogen
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    G = build_generator()
    sd = G.state_dict()
    _, normal, depth = synth_inputs(n_images, H, W, seed=0)
    ogen.generator_forward(sd, normal, depth, use_nlspn_refine=True, nlspn_configs=NLSPN_CFG)      # warm-up
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        ogen.generator_forward(sd, normal, depth, use_nlspn_refine=True, nlspn_configs=NLSPN_CFG)
        ts.append(time.perf_counter() - t0)
    best = min(ts)
    return n_images / best, cores, ts


def run_reference(args, rank, world):
    if rank != 0:
        return
    n = 4
    t_all = []
    import torch
    from _synth import synth_inputs
    from oracle import generator as ogen
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = build_generator().state_dict()
    _, normal, depth = synth_inputs(n, H, W, seed=0)
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        ogen.generator_forward(sd, normal, depth, use_nlspn_refine=True, nlspn_configs=NLSPN_CFG)
        if i >= args.warmup:
            t_all.append(time.perf_counter() - t0)
    val = n * len(t_all) / sum(t_all)
    sample = f"{n} images per step (of the {args.batch}-image batch), {args.steps} steps, fp32, torch CPU + C DCN oracle"
    print(json.dumps({
        "impl": "reference", "metric": "depth maps/sec @228x304", "value": val, "unit": "maps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(t_all) / len(t_all), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, world),
        "cpu_baseline": {"value": val, "unit": "maps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "maps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def workload_config(args, world):
    return {"workload": f"RDFC-GAN RDFGenerator forward (ResNet-18 x2, W-AdaIN, NLSPN TGASS 18 it.), {args.batch} images/GPU "
                        f"@{H}x{W} (BASELINE config 2 batch size = config 3's per-GPU shard at 8 GPUs)",
            "batch_per_gpu": args.batch, "global_batch": args.batch * world, "height": H, "width": W,
            "weights": "synthetic: init_weights recipe + NLSPN stress offsets (tests/_synth.py)",
            "cache": "L2 flushed (256 MiB write) between timed steps; per-step working set >> 126 MB L2",
            "parallelism": f"batch-sharded x{world}, no collective"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    from _synth import synth_inputs
    from rdfc_gan_b200 import _cabi as C
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback exists)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    B = args.batch
    G = build_generator().to(dev).set_precision(args.precision)
    rgb, normal, depth = synth_inputs(B, H, W, seed=rank)
    rgb_d, normal_d, depth_d = rgb.to(dev), normal.to(dev), depth.to(dev)
    with torch.no_grad():
        out = G(rgb_d, depth_d, normal_d)                    # builds + captures the plan
    plan = next(iter(G.engine()._plans.values()))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- value: inputs resident, graph replay, per-step CUDA events, L2 flush between steps
    for _ in range(args.warmup):
        plan.run()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.15)                         # let nvidia-smi start streaming; the GPU keeps running warm-up work meanwhile
    for _ in range(2):
        plan.run()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in evs:
        flush.zero_()
        a.record()
        plan.run()
        b.record()
    barrier()
    clocks = sampler.summary()
    t_ms = sum(a.elapsed_time(b) for a, b in evs)
    tt = torch.tensor([t_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_ms = float(tt.item())
    value = B * world * args.steps / (t_ms / 1e3)

    # ---------------- e2e: public API, pinned host inputs, H2D + D2H inside the timed region
    rgb_h, normal_h, depth_h = rgb.pin_memory(), normal.pin_memory(), depth.pin_memory()
    pred_h = torch.empty(B, 1, H, W).pin_memory()

    def host_batches(n):
        for _ in range(n):
            yield rgb_h, depth_h, normal_h

    def e2e_run(n):
        # public API: G.stream() overlaps the H2D copy of batch i+1 and the D2H read of batch i-1 with the forward of i
        for o in G.stream(host_batches(n), outputs=("pred_depth",)):
            pred_h.copy_(o["pred_depth"])                 # host-side use of the result (pinned -> pinned)
    e2e_run(3)
    barrier()
    t0 = time.perf_counter()
    e2e_run(args.steps)
    barrier()
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e = B * world * args.steps / float(te.item())
    # RDFGenerator's stems read `normal` and `depth` only (rdf_generator.py:286-292: `rgb` is unused), so those are the
    # tensors stream() copies
    h2d = normal_h.numel() * 4 + depth_h.numel() * 4
    d2h = pred_h.numel() * 4

    # ---------------- roofline: NLSPN propagation kernel alone (18 launches per image group)
    T = NLSPN_CFG["prop_time"]
    P = H * W
    group = B          # one launch per iteration over the whole batch (see nlspn.cu)
    n_groups = (B + group - 1) // group
    s = C.stream_ptr()

    def prop_eager():
        C.check(C.lib.rdfc_nlspn_propagate_forward(C.ptr(plan.pred_init), C.ptr(plan.offset), C.ptr(plan.aff), None, 0,
                                                   C.ptr(plan.d2raw), C.ptr(plan.scratch), None, B, H, W, T, 0, C.stream_ptr()))
    for _ in range(3):
        prop_eager()
    torch.cuda.synchronize()
    # the T launches as the product issues them: nodes of a CUDA graph (no host launch gaps between the iterations)
    prop_graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(prop_graph):
        prop_eager()
    prop = prop_graph.replay
    prop()
    torch.cuda.synchronize()
    reps = 10
    pe = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in pe:
        flush.zero_()
        a.record()
        prop()
        b.record()
    torch.cuda.synchronize()
    prop_ms = statistics.median(a.elapsed_time(b) for a, b in pe)
    launches = T * n_groups
    launch_us = prop_ms * 1e3 / launches
    achieved = ALGO_BYTES_PER_PIXEL_ITER * B * P * T / (prop_ms / 1e3) / 1e9
    pk = peaks()
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "nlspn_prop_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
    roofline = {"kernel": "nlspn_prop_band_kernel", "bound": "hbm", "achieved": achieved, "peak": pk["hbm"], "unit": "GB/s",
                "frac": achieved / pk["hbm"], "traffic": traffic, "peak_source": pk["src"] + " (burst copy)",
                "launch_us": launch_us, "launches_per_forward": launches,
                "algorithmic_bytes_per_launch": ALGO_BYTES_PER_PIXEL_ITER * P * min(group, B)}
    dense_tflops = GFLOP_PER_IMAGE * 1e9 * B * args.steps / (t_ms / 1e3) / 1e12
    roofline_dense = {"kernel": "conv_umma_kernel (whole dense part, upper bound: step time includes NLSPN/norm kernels)",
                      "bound": "tensor", "achieved": dense_tflops, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                      "frac": dense_tflops / pk["bf16_sustained"], "peak_source": pk["src"] + " (sustained cuBLAS bf16)"}

    # ---------------- CPU baseline (rank 0, N = 1 only): bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores, ts = cpu_baseline(8, 4)
        cpu = {"value": v, "unit": "maps/s", "cores": cores, "kind": "port",
               "sample": f"8-image batch x 4 repetitions (best) of the same forward (1/4 of the 32-image step), fp32, torch CPU + C DCN oracle; {sum(ts):.1f} s timed"}

    if rank == 0:
        print(json.dumps({
            "metric": "depth maps/sec @228x304", "value": value, "unit": "maps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.precision if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": workload_config(args, world), "clocks": clocks,
            "e2e": {"value": e2e, "unit": "maps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": plan.n_launch * args.steps, "roofline": roofline, "roofline_dense": roofline_dense,
            "cpu_baseline": cpu}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
