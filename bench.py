#!/usr/bin/env python
"""bench.py -- train rays/s (fwd + bwd + Adam) of the PC-NeRF ray-rendering hot path on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--precision tc|fp32]

One "step" = one pass of the hot path over one batch of synthetic LiDAR returns:
    K1 AABB stage (point -> child box, child near/far, parent far, 15-column ray records)
 -> K2 coarse sample placement + positional encoding -> K3 occupancy MLP (coarse net) -> K4 compositing + losses
 -> K2' hierarchical resampling + encoding -> K3 (fine net) -> K4 -> six-term loss (train_kitti.py:145-155)
 -> backward of all of it -> (N > 1: one NCCL all-reduce of the flat gradient buffer) -> Adam step.
Workload at every N (weak scaling): BASELINE.json configs[1] per GPU -- 32,768 rays x 64 coarse samples (+128
importance samples, render_rays_train's own default), ~200 child AABBs, shipped training flags
(segmented sampling ratio 0.1, child losses on, perturb 1, noise_std 0, chunk 262,144).

`value`  : rays/s with the raw returns already resident in HBM.
`e2e`    : the same step driven from pinned HOST buffers (H2D of the returns, D2H of the loss) every step.
`roofline`: the dominant kernel (forward row GEMM of the layered training MLP) against the roofline that bounds it -- HBM at
its 124 FLOP per algorithmic byte (machine balance 210) -- with the tensor-pipe view in `roofline.tensor`; `kernels`: device
time per kernel class and step, with the HBM fraction of every class that has a byte model (DESIGN.md sections 4, 5).
`--impl reference`: the CPU restatement of the reference path (oracle/, kind "port") on a bounded sample of the same
workload with all host threads -- the reference itself is Python that imports from /root/reference and cannot
travel to the GPU box.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

S, NI, K_BOXES, CHUNK = 64, 128, 200, 262144
LAM = (1.0, 1e6, 1e5)
METRIC = "train rays/s (fwd+bwd)"
WORKLOAD = "C2 KITTI-00-shaped block: 32768 rays/GPU x (64 coarse + 128 importance) samples, 200 child AABBs"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default=os.environ.get("PCNERF_PRECISION", "tc"), choices=["fp32", "tc", "affine"],
                    help="MLP engine: tc = TMA + tcgen05 + TMEM (fp16 operands forward, bf16 gradients, fp32 "
                         "accumulation; 1e-3 parity gate), fp32 = CUDA-core SGEMM (1e-5 gate)")
    ap.add_argument("--rays", type=int, default=32768, help="rays per GPU per step")
    ap.add_argument("--cpu-rays", type=int, default=512, help="rays of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--no-inference", action="store_true", help="skip the depth-inference (C3) block")
    ap.add_argument("--no-fast-mode", action="store_true", help="skip the extra closed-form (affine) measurement")
    ap.add_argument("--infer-rays", type=int, default=131072, help="physical LiDAR rays of the inference frame per GPU")
    ap.add_argument("--no-c4", action="store_true", help="skip the strong-scaling C4 block (262,144 rays x 128+256 samples)")
    ap.add_argument("--c4-rays", type=int, default=262144, help="global rays per optimizer step of the C4 block")
    ap.add_argument("--no-c5", action="store_true", help="skip the multi-parent scene block (C5)")
    ap.add_argument("--c5-frames", type=int, default=0, help="frames of the C5 block (0 = a bounded sample: 2 frames per GPU)")
    ap.add_argument("--c5-parents", type=int, default=64)
    ap.add_argument("--graph", default="auto", choices=["auto", "on", "off"],
                    help="replay each step as one CUDA graph (pcnerf_b200.graphed.GraphedStep); auto = fall back to eager "
                         "launches if capture fails")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------- synthetic workload


def make_inputs(rank, n):
    from pcnerf_b200 import synth
    scene = synth.make_scene(1000 + rank, K_BOXES, synth.KITTI_PARENT)
    pts = synth.make_points(scene, 17 + rank, n)
    dirs, dist = synth.rays_from_points(scene.origin, pts)
    return scene, pts, dirs, dist


# ------------------------------------------------------------------------------------------------------- CPU baseline


def cpu_step_fn(n_rays):
    """Oracle port of one step on the host (numpy fp64 AABB stage + torch-CPU renderer), all host threads."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pcnerf_oracle as orc
    torch.set_num_threads(os.cpu_count() or 1)
    scene, pts, dirs, dist = make_inputs(0, n_rays)
    sd_c, sd_f = orc.init_state_dict(42), orc.init_state_dict(43)
    params = []
    for sd in (sd_c, sd_f):
        for k in orc.param_names():
            sd[k].requires_grad_(True)
            params.append(sd[k])
    opt = torch.optim.Adam(params, lr=5e-4, eps=1e-8, weight_decay=1e-3)
    gen = torch.Generator().manual_seed(7)

    def step():
        rays, _ = orc.pack_train_rays_from_dirs(scene.origin, dirs, dist, pts, scene.centres, scene.child_bounds,
                                                scene.child_bounds_bigger, scene.parent, 0.05, "kitti")
        rays = torch.from_numpy(rays)
        n = rays.shape[0]
        U = torch.rand(n, S, generator=gen)
        u = torch.rand(n, NI, generator=gen)
        res = orc.render_rays_train(sd_c, sd_f, rays, S, NI, 1.0, 0, CHUNK, 1, 0.1, 0, 1, U=U, u_fine=u)
        loss = orc.training_loss(res, rays[:, 14], *LAM)
        opt.zero_grad()
        loss.backward()
        opt.step()
        return n, float(loss.detach())

    return step


def reference_step_fn(n_rays):
    """One step of the UNMODIFIED reference (imported from /root/reference through oracle/ref_shim.py): its own
    render_rays_train + the six-term loss of train_kitti.py:145-155 + backward + torch.optim.Adam, on the host cores.  Only
    where the reference tree exists (the build container); the GPU box falls back to the oracle port (cpu_step_fn)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pcnerf_oracle as orc
    import ref_shim
    ref = ref_shim.import_reference()
    torch.set_num_threads(os.cpu_count() or 1)
    scene, pts, dirs, dist = make_inputs(0, n_rays)
    mc, mf, emb = ref.networks.NOF_coarse(), ref.networks.NOF_fine(), ref.networks.Embedding(3, 10)
    mc.load_state_dict(orc.init_state_dict(42))
    mf.load_state_dict(orc.init_state_dict(43))
    mc.train()
    mf.train()
    opt = torch.optim.Adam(list(mc.parameters()) + list(mf.parameters()), lr=5e-4, eps=1e-8, weight_decay=1e-3)
    sl1 = torch.nn.SmoothL1Loss(reduction="mean")

    def step():
        # (the AABB stage of the reference is inlined in its file-reading dataset classes: the oracle restates that loop)
        rays, _ = orc.pack_train_rays_from_dirs(scene.origin, dirs, dist, pts, scene.centres, scene.child_bounds,
                                                scene.child_bounds_bigger, scene.parent, 0.05, "kitti")
        rays = torch.from_numpy(rays)
        with ref_shim.cuda0_to_cpu():
            res = ref.render.render_rays_train(mc, mf, emb, rays, N_samples=S, N_importance=NI, perturb=1.0, noise_std=0,
                                               chunk=CHUNK, issegmentated=1, childnerf_ratio=0.1, use_child_nerf_divide=0,
                                               use_child_nerf_loss=1)
        gt = rays[:, 14]
        loss = 0.1 * LAM[0] * sl1(10 * res["depth"], 10 * gt) + 0.1 * LAM[0] * sl1(10 * res["depth_fine"], 10 * gt) \
            + LAM[1] * (res["child_free_loss_fine"] + res["child_free_loss"]) \
            + LAM[2] * (res["child_depth_loss_fine"] + res["child_depth_loss"])
        opt.zero_grad()
        loss.backward()
        opt.step()
        return rays.shape[0], float(loss.detach())

    return step


def reference_kind():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    try:
        import ref_shim
        return "reference" if ref_shim.reference_available() else "port"
    except Exception:                                          # noqa: BLE001
        return "port"


def time_cpu(n_rays, steps, warmup, kind="port"):
    step = reference_step_fn(n_rays) if kind == "reference" else cpu_step_fn(n_rays)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    rays = 0
    for _ in range(steps):
        n, _ = step()
        rays += n
    dt = time.perf_counter() - t0
    return rays / dt, dt / steps * 1e3


# ------------------------------------------------------------------------------------------------------------ clocks


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0=None, t1=None):
        """Summarise the samples that arrived in [t0, t1] (the timed region); all samples if none fall inside."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 is None or (t0 <= t <= t1 + 0.15)]
        if not rows:
            rows = [r for _, r in self.rows]
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = sorted(sm)[len(sm) // 2:] if sm else []           # upper half = samples under load
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# -------------------------------------------------------------------------------------------------------- the B200 arm


def run_b200(a):
    import torch.distributed as dist
    from pcnerf_b200 import ops, parallel
    from pcnerf_b200.nof import render
    from pcnerf_b200.nof.networks import Embedding, NOF_coarse, NOF_fine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if a.gpus != world and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (a.gpus, world))
    if a.gpus > 1 and world == 1:
        raise SystemExit("--gpus %d needs one process per GPU: launch with python -m torch.distributed.run "
                         "--nproc-per-node %d ..." % (a.gpus, a.gpus))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU baseline")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    n = a.rays
    # Synthetic returns: draw 3 % more than needed and keep the first n that the AABB stage itself accepts (a fraction of a
    # percent of random points misses every child box and would be dropped, ipb2dmapping.py:372-376), so that every timed
    # step packs exactly n rays -- a fixed shape, which is what makes the step capturable as one CUDA graph.
    scene, pts, dirs, dist_v = make_inputs(rank, n + n // 32 + 64)
    _, keep0 = ops.aabb_pack_train(606, scene.origin, dirs, dist_v, pts, scene.centres, scene.child_bounds,
                                   scene.child_bounds_bigger, scene.parent, 0.05, 10, compact=False)
    sel = np.nonzero(keep0.cpu().numpy())[0][:n]
    if sel.shape[0] < n:
        raise SystemExit("synthetic scene produced only %d usable returns" % sel.shape[0])
    pts, dirs, dist_v = pts[sel], dirs[sel], dist_v[sel]
    # scene constants live on the device; the per-step inputs (the LiDAR returns of this batch) exist twice:
    # pinned host buffers (e2e) and device-resident copies (value)
    f64 = dict(dtype=torch.float64, device=dev)
    origin = torch.tensor(scene.origin, **f64)
    centres = torch.tensor(scene.centres, **f64)
    boxes = torch.tensor(scene.child_bounds, **f64)
    boxes_big = torch.tensor(scene.child_bounds_bigger, **f64)
    host = [torch.from_numpy(np.ascontiguousarray(x)).pin_memory() for x in (dirs, dist_v, pts)]
    resident = [h.to(dev) for h in host]
    h2d_bytes = sum(h.numel() * h.element_size() for h in host)

    torch.manual_seed(42 + rank)
    mc, mf = NOF_coarse().to(dev).train(), NOF_fine().to(dev).train()
    if world > 1:                                   # same initial weights on every rank
        for p in list(mc.parameters()) + list(mf.parameters()):
            dist.broadcast(p.data, 0)
    mc.precision = mf.precision = a.precision
    emb = Embedding(3, 10)
    params = list(mc.parameters()) + list(mf.parameters())
    use_graph = a.graph != "off"
    # torch.optim.Adam as the reference configures it, on one flat buffer: one kernel per step, 1/world folded in
    from pcnerf_b200.optim import FlatAdam
    opt = FlatAdam(params, lr=5e-4, eps=1e-8, weight_decay=1e-3)
    bucket = opt.bucket
    sl1 = torch.nn.SmoothL1Loss(reduction="mean")
    dscale = parallel.depth_loss_scale()
    loss_host = torch.zeros(1, dtype=torch.float32).pin_memory()
    staged = [torch.empty_like(r) for r in resident]          # device landing buffers of the e2e graph

    def core(d_, r_, p_, compact, n_s=S, n_i=NI, f_mean=1.0, f_depth=None, zero=True):
        """K1 packing -> render -> six-term loss -> backward into the flat gradient bucket.  (f_mean, f_depth): loss factors
        of a ray micro-batch (pcnerf_b200.train_kitti.microbatch_scales); default = one batch per step per rank."""
        rays, keep = ops.aabb_pack_train(606, origin, d_, r_, p_, centres, boxes, boxes_big, scene.parent, 0.05, 10,
                                         compact=compact)
        res = render.render_rays_train(mc, mf, emb, rays, N_samples=n_s, N_importance=n_i, perturb=1.0, noise_std=0,
                                       chunk=CHUNK, issegmentated=1, childnerf_ratio=0.1, use_child_nerf_divide=0,
                                       use_child_nerf_loss=1)
        f_depth = dscale if f_depth is None else f_depth
        # (the two scene-level range terms SmoothL1(10 depth, 10 gt) come out of K4's compositing pass: res["range_sl1*"])
        loss = f_mean * (0.1 * LAM[0] * (res["range_sl1"] + res["range_sl1_fine"])
                         + LAM[1] * (res["child_free_loss_fine"] + res["child_free_loss"])) \
            + LAM[2] * f_depth * (res["child_depth_loss_fine"] + res["child_depth_loss"])
        if zero:
            bucket.zero()
        loss.backward()
        return loss.detach().reshape(1), rays.shape[0], keep

    def finish():
        """Gradient all-reduce (N > 1) + Adam.  Kept OUT of the captured graph when N > 1: NCCL collectives replayed from a
        CUDA graph hung the process at teardown on this stack (profiles/README.md), and the step has exactly one."""
        opt.step()                                   # (all-reduce of the flat gradients when N > 1, then the update)

    def step(from_host):
        """One eager step (every kernel launched from the host)."""
        if from_host:
            d_, r_, p_ = [h.to(dev, non_blocking=True) for h in host]
        else:
            d_, r_, p_ = resident
        loss, nrays, _ = core(d_, r_, p_, True)
        finish()
        if from_host:
            loss_host.copy_(loss, non_blocking=True)
            torch.cuda.current_stream().synchronize()          # the caller reads the loss every step
        return nrays

    def build_graphs():
        """Capture the resident-input and the host-input step as CUDA graphs; ({}, note) when graphs are off or fail."""
        if not use_graph:
            return {}, "off"
        try:
            from pcnerf_b200.graphed import GraphedStep
            _, _, keep = core(*resident, True)
            finish()
            if not bool(keep.all()):
                raise RuntimeError("the AABB stage drops rays of this batch: data-dependent shape, not capturable")

            def g_resident():
                core(*resident, False)
                if world == 1:
                    finish()

            def g_host():
                for s_, h_ in zip(staged, host):
                    s_.copy_(h_, non_blocking=True)
                loss, _, _ = core(*staged, False)
                if world == 1:
                    finish()
                loss_host.copy_(loss, non_blocking=True)

            return {"resident": GraphedStep(g_resident, warmup=2), "host": GraphedStep(g_host, warmup=1)}, "on"
        except Exception as exc:                               # noqa: BLE001 - any capture failure -> eager launches
            if a.graph == "on":
                raise
            import traceback
            sys.stderr.write("bench.py: CUDA-graph capture failed (%s: %s); falling back to eager launches\n%s\n"
                             % (type(exc).__name__, exc, "".join(traceback.format_tb(exc.__traceback__)[-6:])))
            torch.cuda.synchronize()
            return {}, "capture failed, eager"

    graphs, graph_note = build_graphs()

    def run(from_host):
        if graphs:
            graphs["host" if from_host else "resident"]()
            if world > 1:
                finish()
            if from_host:
                torch.cuda.current_stream().synchronize()
            return n
        return step(from_host)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(from_host, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ops.launch_count(reset=True)
        e0.record()
        rays_done = 0
        for _ in range(steps):
            rays_done += run(from_host)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        tot = torch.tensor([float(rays_done)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        return float(ms.item()), float(tot.item()), ops.launch_count()

    def time_c4():
        """BASELINE.json configs[3]: ONE optimizer step over 262,144 rays x (128 coarse + 256 importance) samples
        (nof/nof_utils.py:121-124 defaults) with the rays STRONG-sharded over the ranks (262,144 / N per GPU), every rank
        accumulating gradients over micro-batches of 32,768 rays (the saved activations of one micro-batch are 69 GB), then
        one NCCL all-reduce of the flat gradient buffer + Adam.  BatchNorm batches are the same 262,144-row chunks the
        reference would form on the whole batch (a micro-batch holds 16 / 48 whole chunks).  value = global rays / s."""
        from pcnerf_b200 import synth
        from pcnerf_b200.graphed import GraphedStep
        from pcnerf_b200.train_kitti import microbatch_scales
        NG, S4, NI4, MB = a.c4_rays, 128, 256, 32768
        n_l = NG // world
        mb = min(MB, n_l)
        nmb = n_l // mb
        pts4 = synth.make_points(scene, 4000 + rank, n_l + n_l // 32 + 64)
        dirs4, dist4 = synth.rays_from_points(scene.origin, pts4)
        _, keep4 = ops.aabb_pack_train(606, scene.origin, dirs4, dist4, pts4, scene.centres, scene.child_bounds,
                                       scene.child_bounds_bigger, scene.parent, 0.05, 10, compact=False)
        sel4 = np.nonzero(keep4.cpu().numpy())[0][:n_l]
        res4 = [torch.from_numpy(np.ascontiguousarray(x[sel4])).to(dev) for x in (dirs4, dist4, pts4)]
        stage4 = [torch.empty_like(r[:mb]) for r in res4]
        f_mean, f_depth = microbatch_scales(mb, n_l, world)

        def micro():
            core(*stage4, False, S4, NI4, f_mean, f_depth, zero=False)

        g4, note4 = None, "off"
        if use_graph:
            try:
                for s_, r_ in zip(stage4, res4):
                    s_.copy_(r_[:mb])
                g4, note4 = GraphedStep(micro, warmup=1), "on (one graph per ray micro-batch)"
            except Exception as exc:                           # noqa: BLE001
                if a.graph == "on":
                    raise
                sys.stderr.write("bench.py: c4 graph capture failed (%s: %s); eager launches\n" % (type(exc).__name__, exc))
                torch.cuda.synchronize()
                g4, note4 = None, "capture failed, eager"

        def step4():
            bucket.zero()
            for m in range(nmb):
                for s_, r_ in zip(stage4, res4):
                    s_.copy_(r_[m * mb:(m + 1) * mb])
                if g4 is not None:
                    g4()
                else:
                    micro()
            finish()

        step4()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k4 = max(1, min(a.steps, 2))
        e0.record()
        for _ in range(k4):
            step4()
        e1.record()
        barrier()
        ms4 = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms4, op=dist.ReduceOp.MAX)
        ms4 = float(ms4.item()) / k4
        del g4
        return {"workload": "C4: %d rays x (128 coarse + 256 importance) samples per optimizer step, rays strong-sharded "
                            "over %d GPU(s), %d micro-batch(es) of %d rays per GPU with gradient accumulation, one flat NCCL "
                            "all-reduce + Adam per step" % (NG, world, nmb, mb),
                "metric": METRIC, "value": NG / (ms4 * 1e-3), "unit": "rays/s", "scaling": "strong", "ms_per_step": ms4,
                "rays_global": NG, "rays_per_gpu": n_l, "micro_batches_per_gpu": nmb, "micro_batch_rays": mb,
                "N_samples": S4, "N_importance": NI4, "chunk": CHUNK, "steps": k4, "warmup": 1, "cuda_graph": note4,
                "sample_evals_per_step": NG * (S4 + S4 + NI4), "precision": a.precision}

    clocks = ClockSampler(local)                     # started before the warm-up: nvidia-smi needs ~1 s to come up
    clocks.start()
    # kernels launched by one step (a graph replay re-issues exactly the launches recorded at capture)
    ops.launch_count(reset=True)
    step(False)
    launches_per_step = ops.launch_count()
    for _ in range(max(a.warmup, 3)):
        run(False)
    run(True)
    t_wall0 = time.time()
    ms, rays_total, _ = timed(False, a.steps)
    launches = launches_per_step * a.steps
    ms_e2e, rays_e2e, _ = timed(True, a.steps)
    clk = clocks.stop(t_wall0, time.time())

    # ---- per-kernel-class device time (separate pass so the event pairs do not perturb the numbers above)
    roofline, kernels = None, None
    if not a.no_profile:
        barrier()
        lanes = ops.TC_LANES
        ops.TC_LANES = 1              # one chunk at a time: event intervals of concurrent kernels would overlap
        ops.profile(True)
        psteps = min(a.steps, 2)
        for _ in range(psteps):
            step(False)
        torch.cuda.synchronize()
        prof = ops.profile_read()
        ops.profile(False)
        ops.TC_LANES = lanes
        kernels = {k: {"ms_per_step": v[0] / psteps, "launches_per_step": v[1] / psteps} for k, v in prof.items() if v[1]}
        gemm = [prof[k] for k in ("mlp_gemm_fwd", "mlp_gemm_dgrad", "mlp_gemm_wgrad")]
        g_ms, g_fl = sum(g[0] for g in gemm), sum(g[2] for g in gemm)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except (OSError, ValueError):
            pass
        # HBM-bound kernel classes: algorithmic bytes (counted by the library per launch) / device time
        for k in ("sample_encode", "composite_fwd", "composite_bwd", "aabb", "search"):
            if k in kernels and prof[k][0] > 0 and prof[k][2] > 0:
                gbs = prof[k][2] / (prof[k][0] * 1e-3) / 1e9
                kernels[k]["algorithmic_gbs"] = gbs
                kernels[k]["hbm_frac"] = gbs / float(peaks.get("hbm_gbs", 6650.0))
        peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        all_ms = max(sum(v[0] for v in prof.values()), 1e-9)
        f_ms, f_n, f_fl = prof["mlp_gemm_fwd"]
        d_ms, d_n, d_fl = prof["mlp_gemm_dgrad"]
        w_ms, w_n, w_fl = prof["mlp_gemm_wgrad"]
        esz = 2 if a.precision == "tc" else 4
        # ALGORITHMIC bytes per launch: forward = one activation matrix in, one out per layer (layer 0
        # reads the (rows,64) encoding, layer 4 both; the correction blocks of k_tc_fold re-read the L2-resident encoding and
        # are overhead, not algorithmic traffic); data gradient = DH_l in, H_{l-1} in, DH_{l-1} out; weight gradient = DH_l and
        # H_{l-1} (or the encoding) in
        by = {"mlp_gemm_fwd": CHUNK * esz * (7 * 256 + 2 * 64 + 8 * 256) / 8.0 * f_n,
              "mlp_gemm_dgrad": d_n * CHUNK * 256.0 * esz * 3,
              "mlp_gemm_wgrad": (w_n / 9.0) * CHUNK * esz * (7 * 512.0 + 2 * 320.0)}
        names = {"mlp_gemm_fwd": "mlp forward row GEMM (%s)" % ("k_tc_rowgemm2<FWD>: TMA + tcgen05.mma.cta_group::2 + TMEM, CTA pairs"
                                                                 if a.precision == "tc" else "k_gemm fp32"),
                 "mlp_gemm_dgrad": "mlp data-gradient row GEMM with the BN backward in its epilogue (%s)"
                                   % ("k_tc_rowgemm2<DGRAD2>: TMA + tcgen05.mma.cta_group::2 + TMEM, CTA pairs, the H term of the "
                                      "BN backward on the tensor core" if a.precision == "tc" else "k_gemm fp32"),
                 "mlp_gemm_wgrad": "mlp weight-gradient GEMM, split-K over the rows (%s)"
                                   % ("k_tc_wgrad2: TMA + tcgen05.mma.cta_group::2 + TMEM, CTA pairs" if a.precision == "tc"
                                      else "k_gemm fp32")}
        cls = {}
        for k, (ms_k, n_k, fl_k) in (("mlp_gemm_fwd", (f_ms, f_n, f_fl)), ("mlp_gemm_dgrad", (d_ms, d_n, d_fl)),
                                    ("mlp_gemm_wgrad", (w_ms, w_n, w_fl))):
            if ms_k > 0 and n_k > 0:
                gbs = by[k] / (ms_k * 1e-3) / 1e9
                tfl = fl_k / (ms_k * 1e-3) / 1e12
                kernels[k]["algorithmic_gbs"] = gbs
                kernels[k]["hbm_frac"] = gbs / hbm_peak
                cls[k] = {"ms_per_step": ms_k / psteps, "us_per_launch": 1e3 * ms_k / n_k, "launches_per_step": n_k / psteps,
                          "algorithmic_gbs": gbs, "hbm_frac": gbs / hbm_peak, "tflops": tfl, "tensor_frac": tfl / peak,
                          "share_of_step": ms_k / all_ms, "algorithmic_bytes_per_launch": by[k] / n_k,
                          "intensity_flop_per_byte": fl_k / by[k]}
        g_ms, g_fl = f_ms + d_ms + w_ms, f_fl + d_fl + w_fl
        gemms = {"achieved": g_fl / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0, "unit": "TFLOP/s",
                 "frac": (g_fl / (g_ms * 1e-3) / 1e12 / peak) if g_ms > 0 else 0.0, "share_of_step": g_ms / all_ms,
                 "hbm_frac": (sum(by.values()) / (g_ms * 1e-3) / 1e9 / hbm_peak) if g_ms > 0 else 0.0}
        traffic, traffic_src = None, None
        dom = max(cls, key=lambda k: cls[k]["ms_per_step"]) if cls else None
        if dom is not None:
            try:
                tr = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
                if a.precision == "tc" and CHUNK == tr["rows"] and dom in tr:
                    traffic = tr[dom]["dram_bytes_read"] + tr[dom]["dram_bytes_write"]
                    traffic_src = "static: one ncu --set full capture of %s (profiles/r02_traffic.json: %s), not measured in this run" % (tr[dom].get("kernel", "this kernel"), tr[dom]["source"])
            except (OSError, ValueError, KeyError):
                pass
            c = cls[dom]
            if a.precision == "tc":
                # 16-bit layered GEMMs: 124-149 FLOP per algorithmic byte against a machine balance of 210 -> HBM is the
                # roofline that bounds every class (train-mode BN forces each layer's activations through HBM once)
                roofline = {"kernel": names[dom], "class": dom, "dominant_by": "largest device time per step of the three GEMM classes",
                            "bound": "hbm", "achieved": c["algorithmic_gbs"], "peak": hbm_peak,
                            "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650", "unit": "GB/s",
                            "frac": c["hbm_frac"], "traffic": traffic, "traffic_source": traffic_src,
                            "algorithmic_bytes_per_launch": c["algorithmic_bytes_per_launch"],
                            "intensity_flop_per_byte": c["intensity_flop_per_byte"],
                            "machine_balance_flop_per_byte": peak * 1e3 / hbm_peak, "us_per_launch": c["us_per_launch"],
                            "launches_per_step": c["launches_per_step"], "share_of_step": c["share_of_step"],
                            "tensor": {"achieved": c["tflops"], "peak": peak, "unit": "TFLOP/s", "frac": c["tensor_frac"],
                                       "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1400",
                                       "attainable_tflops": c["intensity_flop_per_byte"] * hbm_peak / 1e3},
                            "classes": cls, "all_mlp_gemms": gemms}
            else:
                roofline = {"kernel": names[dom], "class": dom, "bound": "tensor", "achieved": c["tflops"], "peak": peak,
                            "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1400",
                            "unit": "TFLOP/s", "frac": c["tensor_frac"], "traffic": None, "us_per_launch": c["us_per_launch"],
                            "launches_per_step": c["launches_per_step"], "share_of_step": c["share_of_step"],
                            "classes": cls, "all_mlp_gemms": gemms}

    out = {"metric": METRIC, "value": rays_total / (ms * 1e-3), "unit": "rays/s", "n_gpus": world, "steps": a.steps,
           "warmup": max(a.warmup, 3), "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": {"fp32": "f32", "tc": "f16 fwd / bf16 bwd operands, f32 accumulate", "affine": "f32 data kernels, f64 closed-form algebra"}[a.precision], "data": "synthetic",
           "config": {"workload": WORKLOAD, "rays_per_gpu": n, "N_samples": S, "N_importance": NI, "chunk": CHUNK,
                      "child_aabbs": K_BOXES, "precision": a.precision, "optimizer": "Adam on one flat buffer (pcnerf_b200.optim.FlatAdam)", "cuda_graph": graph_note,
                      "mlp_chunks_in_flight": ops.TC_LANES if a.precision == "tc" else 1,
                      "parallelism": "dp%d (rays sharded, one flat NCCL all-reduce of 3.98 MB grads)" % world,
                      "l2": "per-step working set (>2 GB encodings, >30 GB activations) exceeds the 126 MB L2"},
           "clocks": clk,
           "e2e": {"value": rays_e2e / (ms_e2e * 1e-3), "unit": "rays/s", "h2d_bytes_per_step": h2d_bytes,
                   "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / a.steps},
           "gpu_launches": launches}
    if roofline:
        out["roofline"] = roofline
    if kernels:
        out["kernels"] = kernels
    if not a.no_inference:
        out["inference"] = time_inference(a, rank, world, dev, mc, mf, emb)
    if a.precision != "affine" and not a.no_fast_mode:
        # The same step with the MLP engine switched to the closed form (precision 2): reported beside
        # the headline, never instead of it.
        graphs.clear()
        torch.cuda.empty_cache()
        mc.precision = mf.precision = "affine"
        g2, note2 = build_graphs()
        graphs.update(g2)
        for _ in range(3):
            run(False)
        run(True)
        ms_f, rays_f, _ = timed(False, a.steps)
        ms_fe, rays_fe, _ = timed(True, a.steps)
        fast = {"precision": "affine (closed form of the identity-activation network, 1e-5 parity gate: "
                             "tests/test_gpu_affine.py)", "value": rays_f / (ms_f * 1e-3), "unit": "rays/s",
                "ms_per_step": ms_f / a.steps, "e2e": rays_fe / (ms_fe * 1e-3), "cuda_graph": note2}
        if not a.no_profile:
            # where the closed-form step goes (library CUDA events per kernel class, eager launches) and the roofline of its
            # dominant kernel: the second-moment kernel is fp32 FMA work (rows x 2,304 FMA, upper triangle of x x^T), bounded
            # by the CUDA-core FMA pipe -- neither HBM (8 bytes per row are read) nor the tensor cores
            barrier()
            clocks_f = ClockSampler(local)                # own clock samples: this step is not power-capped like the headline
            clocks_f.start()
            time.sleep(1.2)                               # nvidia-smi needs ~1 s to come up
            t_f0 = time.time()
            ops.profile(True)
            psteps = 60                                   # ~0.25 s of eager steps: a few 100-ms clock samples under load
            for _ in range(psteps):
                step(False)
            torch.cuda.synchronize()
            prof_f = ops.profile_read()
            ops.profile(False)
            clk_f = clocks_f.stop(t_f0, time.time())
            fast["clocks"] = clk_f
            fast["kernels"] = {k: {"ms_per_step": v[0] / psteps, "launches_per_step": v[1] / psteps}
                               for k, v in prof_f.items() if v[1]}
            m_ms, m_n, m_fl = prof_f.get("affine_moments", (0.0, 0, 0.0))
            a_ms, a_n, a_fl = prof_f.get("affine_algebra", (0.0, 0, 0.0))
            sm_mhz = float((clk_f or {}).get("sm_mhz") or 0.0) or float((clk_f or {}).get("sm_max_mhz") or 0.0) or 1965.0
            fma_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
            if m_ms > 0:
                tf = m_fl / (m_ms * 1e-3) / 1e12
                fast["roofline"] = {"kernel": "k_affine_moments_rays (per-chunk second moments of the re-derived encodings: "
                                              "8 x 8 register tiles, packed fp32 FFMA2, fp64 fold)",
                                    "bound": "fp32 FMA pipe (CUDA cores)", "achieved": tf, "peak": fma_peak, "unit": "TFLOP/s",
                                    "peak_source": "148 SMs x 128 FMA lanes x 2 FLOP x %.0f MHz (median SM clock sampled by this "
                                                   "run)" % sm_mhz, "frac": tf / fma_peak,
                                    "share_of_step": m_ms / max(sum(v[0] for v in prof_f.values()), 1e-9),
                                    "ms_per_step": m_ms / psteps,
                                    "algorithmic_flop_per_row": 36 * 64 * 2, "hbm_bytes_per_row": 8,
                                    "note": "encoding re-derivation (30 sin/cos pairs per row, ~24 instructions each) runs on "
                                            "the same pipes inside this kernel and is not counted as algorithmic work"}
            if a_ms > 0:
                fast["algebra"] = {"what": "float64 parameter-sized chain (mma.sync m8n8k4 f64 GEMMs + per-layer kernels), "
                                           "forward and hand-derived backward", "ms_per_step": a_ms / psteps,
                                   "tflops_f64": a_fl / (a_ms * 1e-3) / 1e12, "launches_per_step": a_n / psteps}
        if not a.no_inference:
            inf_f = time_inference(a, rank, world, dev, mc, mf, emb)
            fast["inference"] = inf_f["value"]
            fast["inference_ms_per_frame"] = inf_f.get("ms_per_frame")
            fast["inference_kernels"] = inf_f.get("kernels")
        out["fast_mode"] = fast
        graphs.clear()
        mc.precision = mf.precision = a.precision
    if not a.no_c4:
        graphs.clear()
        torch.cuda.empty_cache()
        out["c4"] = time_c4()
    if not a.no_c5:
        torch.cuda.empty_cache()
        out["c5"] = time_c5(a, rank, world, dev, emb)
    # the driver keeps `config` of every per-N line: compact copies of the other workloads' headline numbers go there
    other = {}
    if "inference" in out:
        other["c3_depth_inference"] = {k: out["inference"][k] for k in ("value", "unit", "ms_per_frame", "physical_rays_per_gpu")}
    if "c4" in out:
        other["c4_train_strong"] = {k: out["c4"][k] for k in ("value", "unit", "ms_per_step", "rays_global", "rays_per_gpu")}
    if "c5" in out:
        other["c5_multi_parent_inference"] = {k: out["c5"][k] for k in ("value", "unit", "frames_per_s", "frames_timed", "parents")}
    out["config"]["other_workloads"] = other
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        v, ms_cpu = time_cpu(a.cpu_rays, 2, 1)
        out["cpu_baseline"] = {"value": v, "unit": "rays/s", "cores": os.cpu_count(), "kind": "port",
                               "sample": "%d rays of the same workload (oracle port, fp32, same S/Ni/chunk/flags), "
                                         "1 warm-up + 2 timed steps, %.0f ms/step" % (a.cpu_rays, ms_cpu)}
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)          # skip interpreter teardown of NCCL + CUDA-graph state (a hang here cost a 15-minute box)


def time_inference(a, rank, world, dev, mc, mf, emb):
    """BASELINE.json configs[2]: one full 64-beam frame (131,072 physical rays/GPU, every rank renders its own frame =
    rays sharded with no communication), two-step parent-then-child search, ~2.9 candidate rows per ray with the shipped
    group-size histogram, 64 + 128 samples, eval-mode BN, through pcnerf_b200.eval_kitti_render.render_frame.
    rays/s counts PHYSICAL rays.  Timed with the candidate rows already resident in HBM."""
    import torch.distributed as dist
    from pcnerf_b200 import eval_kitti_render as ev
    from pcnerf_b200 import ops as ops_mod
    from pcnerf_b200 import synth
    base_phys = 4096
    rows, other, true_range = synth.synth_infer_rows(500 + rank, base_phys)
    reps = max(1, a.infer_rays // base_phys)
    rows = np.tile(rows, (reps, 1))
    other = np.tile(other, reps)
    n_phys = base_phys * reps
    rays_d = torch.from_numpy(rows).to(dev)
    other_d = torch.from_numpy(other).to(dev)
    mc.eval()
    mf.eval()

    # group structure of the frame's candidate rows (which rows start a physical ray): produced once with the rows -- by the
    # group builder K1 (ops.aabb_build_groups) or when a cached frame is loaded -- not per rendering
    plan = ev.FramePlan(rays_d, other_d, 18432)

    def frame():
        return ev.render_frame(mc, mf, emb, rays_d, other_d, S, NI, 184320, depth_inference_method=2, batch_size_set=18432,
                               plan=plan)

    pts = frame()
    assert pts.shape[0] == n_phys, (pts.shape, n_phys)          # exactly one rendered point per physical ray
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    reps_t = 3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps_t):
        frame()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps_t], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    # where the frame goes (separate pass with the library's per-class CUDA events) and the tensor roofline of the
    # eval-mode MLP kernel (k_tc_fused_eval when precision = tc): 982,528 FLOP per sample evaluation
    ops_mod.profile(True)
    frame()
    torch.cuda.synchronize()
    prof = ops_mod.profile_read()
    ops_mod.profile(False)
    classes = {k: {"ms_per_frame": v[0], "launches": v[1]} for k, v in prof.items() if v[1]}
    mlp_roof = None
    f_ms, f_n, f_fl = prof["mlp_gemm_fwd"]
    if f_ms > 0 and a.precision == "tc":
        try:
            peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"])
            src = "MEASURED_PEAKS.json bf16_tflops_sustained"
        except (OSError, ValueError, KeyError):
            peak, src = 1400.0, "fallback 1400"
        tf = f_fl / (f_ms * 1e-3) / 1e12
        mlp_roof = {"kernel": "k_tc_fused_eval (eval-mode MLP, all layers in one tcgen05 kernel, CTA pairs)", "bound": "tensor",
                    "achieved": tf, "peak": peak, "peak_source": src, "unit": "TFLOP/s", "frac": tf / peak,
                    "launches_per_frame": f_n, "share_of_frame": f_ms / float(ms.item())}
    # downstream step: Chamfer distance / F-score of the rendered cloud against the returns
    from pcnerf_b200.nof.criteria.metrics import eval_points
    heads = np.nonzero(rows[:base_phys * 40, 12] >= 0)[0][:base_phys]
    gt = torch.from_numpy(np.tile(rows[heads, :3] + rows[heads, 3:6] * true_range[:, None].astype(np.float32), (reps, 1))).to(dev)
    eval_points(pts, gt)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    cd, fscore = eval_points(pts, gt)
    torch.cuda.synchronize()
    metrics_ms = (time.perf_counter() - t0) * 1e3
    # upstream step (K1, eval_kitti_render.py:353-461): candidate-group construction for one frame of raw rays against
    # ~200 child AABBs (parent slab test, 0.65 m prefilter, exactly-two-hits intersection, grow-until-hit fallback, sort)
    scene = synth.make_scene(900 + rank, K_BOXES, synth.KITTI_PARENT)
    fp = synth.make_points(scene, 5 + rank, n_phys)
    fd, fr = synth.rays_from_points(scene.origin, fp)
    sbl = scene.child_bounds + np.array([-0.025] * 3 + [0.025] * 3)
    f64 = dict(dtype=torch.float64, device=dev)
    g_args = (torch.tensor(scene.origin, **f64), torch.tensor(fd, **f64), torch.tensor(fr, **f64),
              torch.tensor(scene.child_bounds, **f64), torch.tensor(sbl, **f64), scene.parent_min, scene.parent_max, 2, 0.05, 0.65)
    g_rows = ops_mod.aabb_build_groups(*g_args)[0]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ops_mod.aabb_build_groups(*g_args)
    torch.cuda.synchronize()
    groups_ms = (time.perf_counter() - t0) * 1e3
    mc.train()
    mf.train()
    return {"metric": "depth-inference rays/s (physical rays, two-step search)", "eval_pts_ms": metrics_ms,
            "aabb_groups_ms": groups_ms, "aabb_groups_rows": int(g_rows.shape[0]),
            "chamfer_untrained_net": cd, "value": world * n_phys / (float(ms.item()) * 1e-3),
            "unit": "rays/s", "ms_per_frame": float(ms.item()), "physical_rays_per_gpu": n_phys,
            "candidate_rows_per_gpu": int(rows.shape[0]), "N_samples": S, "N_importance": NI, "batch_rows": 18432,
            "precision": a.precision, "kernels": classes, "roofline": mlp_roof}


def time_c5(a, rank, world, dev, emb):
    """BASELINE.json configs[4]: a large scene of 64 parent blocks (8 x 8 grid of 40 x 40 x 2.2 m KITTI-like blocks), each
    with ~500 child AABBs and its OWN pair of occupancy networks, the blocks sharded over the GPUs (8 per GPU at 8 GPUs);
    batched depth inference of full 64-beam frames (131,072 returns each; the sensor of frame f stands in block f mod 64 and
    sees that block and its neighbours): returns routed to their block (K0'), candidate groups per block over all frames of
    the batch (K1), two-step search rendering once per physical ray.  No communication.  value = physical rays / s over all
    GPUs; frames_per_s = value / 131,072.  A default run times a bounded sample (2 frames per GPU per batch, the batch
    rendered `reps` times); --c5-frames 1000 renders 1,000 frames."""
    import torch.distributed as dist
    from pcnerf_b200 import scene as sc
    from pcnerf_b200 import synth
    from pcnerf_b200.nof.networks import NOF_coarse, NOF_fine
    P, K5, RAYS = a.c5_parents, 500, 131072
    side = int(round(P ** 0.5))
    assert side * side == P, "--c5-parents must be a square number"
    bw = 40.0
    rng = np.random.default_rng(5000)
    parents, scenes = [], []
    # adjacent blocks on different ranks (scene.spread_owners): a frame only sees the 3 x 3 blocks around its sensor
    owners = [((i % side) + 3 * (i // side)) % world for i in range(P)]
    owned = set(sc.owned_blocks(P, world, rank, owners))
    for i in range(P):
        gx, gy = i % side, i // side
        box = (gx * bw, (gx + 1) * bw, gy * bw, (gy + 1) * bw, -1.7, 0.5)
        s_i = synth.make_scene(7000 + i, K5, box)
        scenes.append(s_i)
        blk = sc.ParentBlock(s_i.parent_min, s_i.parent_max, s_i.child_bounds,
                             s_i.child_bounds + np.array([-0.025] * 3 + [0.025] * 3))
        if i in owned:
            torch.manual_seed(9000 + i)
            blk.nof_coarse, blk.nof_fine = NOF_coarse().to(dev).eval(), NOF_fine().to(dev).eval()
            blk.nof_coarse.precision = blk.nof_fine.precision = a.precision
        parents.append(blk)
    # one batch of frames (every rank builds the same batch: frames are broadcast, blocks are sharded)
    fb = 2 * world
    origins, pts, fid = [], [], []
    for f in range(fb):
        # the frames of a batch are spread over the whole scene (a batch of a 1,000-frame job spans the trajectory), so that
        # every rank's parent blocks receive rays: sensor blocks spaced P / fb apart
        gx, gy = (3 * f + 1) % side, ((f * side) // fb + 1) % side
        nb = [(x, y) for x in range(max(gx - 1, 0), min(gx + 2, side)) for y in range(max(gy - 1, 0), min(gy + 2, side))]
        o = np.array([(gx + 0.5) * bw + rng.uniform(-5, 5), (gy + 0.5) * bw + rng.uniform(-5, 5), rng.uniform(-0.3, 0.0)])
        per = RAYS // len(nb)
        for j, (x, y) in enumerate(nb):
            n_j = per if j < len(nb) - 1 else RAYS - per * (len(nb) - 1)
            pts.append(synth.make_points(scenes[y * side + x], 100 * f + j, n_j))
            fid.append(np.full(n_j, f, dtype=np.int32))
        origins.append(o)
    origins_d = torch.tensor(np.stack(origins), dtype=torch.float64, device=dev)
    pts_d = torch.tensor(np.concatenate(pts), dtype=torch.float64, device=dev)
    fid_d = torch.tensor(np.concatenate(fid), dtype=torch.int32, device=dev)
    boxes_d = torch.as_tensor(sc.parent_boxes(parents), dtype=torch.float64, device=dev)

    def batch():
        res = sc.render_scene_frames(parents, origins_d, pts_d, fid_d, emb, S, NI, 184320, 2, 18432, 0.05, world, rank,
                                     boxes_dev=boxes_d, owners=owners)
        return sum(int(v.shape[0]) for v in res.values())

    rendered = batch()                                     # warm-up (also loads every block's folded weights)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    frames_total = a.c5_frames if a.c5_frames > 0 else 2 * fb
    reps = max(1, -(-frames_total // fb))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        batch()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    tot = torch.tensor([float(rendered)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    sec = float(ms.item()) * 1e-3
    frames = reps * fb
    return {"workload": "C5: %d parent blocks x %d child AABBs, own networks per block, blocks sharded over %d GPU(s), batched "
                        "two-step depth inference, %d frames x %d returns per batch" % (P, K5, world, fb, RAYS),
            "metric": "depth-inference rays/s (physical rays, multi-parent scene)", "value": frames * RAYS / sec,
            "unit": "rays/s", "frames_per_s": frames / sec, "frames_timed": frames, "frames_per_batch": fb,
            "parents": P, "parents_per_gpu": len(owned), "block_ownership": "(ix + 3 iy) mod N: adjacent blocks on different ranks", "child_aabbs_per_parent": K5,
            "rendered_points_per_batch": float(tot.item()), "returns_per_batch": fb * RAYS, "ms_per_batch": 1e3 * sec / reps,
            "N_samples": S, "N_importance": NI, "precision": a.precision, "scaling": "weak (frames per batch grow with N)"}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kind = reference_kind()
    v, ms = time_cpu(a.cpu_rays, a.steps, a.warmup, kind)
    what = "the unmodified reference (imported from /root/reference)" if kind == "reference" else \
        "oracle port of the reference's CPU path (the reference tree does not travel to the GPU box)"
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": "rays/s", "n_gpus": a.gpus, "steps": a.steps,
           "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic",
           "config": {"workload": WORKLOAD, "sample_rays_per_step": a.cpu_rays, "N_samples": S, "N_importance": NI,
                      "chunk": CHUNK, "child_aabbs": K_BOXES},
           "cpu_baseline": {"value": v, "unit": "rays/s", "cores": os.cpu_count(), "kind": kind,
                            "sample": "%d rays/step of the same workload, %s" % (a.cpu_rays, what)},
           "e2e": {"value": v, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
