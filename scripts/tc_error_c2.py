#!/usr/bin/env python
"""GPU: relative depth error of the tensor-core engine against the fp32 engine on one C2 training pass (32,768 rays x
64 + 128 samples), with the linear weight-rounding correction (k_tc_fold) off and on, and the forward/backward time of
both forms.  Prints one JSON line."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from gpu_util import dev, make_nets  # noqa: E402
from pcnerf_b200 import ops, synth  # noqa: E402
from pcnerf_b200.nof import render  # noqa: E402

N, S, NI, CHUNK = int(os.environ.get("RAYS", 32768)), 64, 128, 262144


def run(prec, rays, U, u):
    mc, mf, emb = make_nets(42, 43, True, prec)
    res = render.render_rays_train(mc, mf, emb, rays, N_samples=S, N_importance=NI, perturb=1.0, noise_std=0, chunk=CHUNK,
                                   issegmentated=1, childnerf_ratio=0.1, use_child_nerf_divide=0, use_child_nerf_loss=1,
                                   U=U, u=u)
    loss = res["depth"].mean() + res["depth_fine"].mean() + 1e6 * (res["child_free_loss"] + res["child_free_loss_fine"])
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    loss.backward()
    torch.cuda.synchronize()
    # timed second pass
    for p in list(mc.parameters()) + list(mf.parameters()):
        p.grad = None
    t0.record()
    res2 = render.render_rays_train(mc, mf, emb, rays, N_samples=S, N_importance=NI, perturb=1.0, noise_std=0, chunk=CHUNK,
                                    issegmentated=1, childnerf_ratio=0.1, use_child_nerf_divide=0, use_child_nerf_loss=1,
                                    U=U, u=u)
    (res2["depth"].mean() + res2["depth_fine"].mean()).backward()
    t1.record()
    torch.cuda.synchronize()
    out = {k: v.detach().float().cpu().numpy() for k, v in res.items()}
    g = mc.layer1[3].weight.grad.cpu().numpy()
    return out, t0.elapsed_time(t1), g


def main():
    rays = torch.from_numpy(synth.synth_train_rays(2024, N, K=200, parent=synth.KITTI_PARENT)).to(dev())
    U = torch.rand((N, S), device=dev(), generator=torch.Generator(device=dev()).manual_seed(4))
    u = torch.rand((N, NI), device=dev(), generator=torch.Generator(device=dev()).manual_seed(5))
    ref, _, gref = run("fp32", rays, U, u)
    out = {"rays": N}
    for corr in (0, 1):
        ops.tc_weight_correction(corr)
        o, ms, g = run("tc", rays, U, u)
        rel = np.abs(o["depth"] - ref["depth"]) / np.abs(ref["depth"])
        relf = np.abs(o["depth_fine"] - ref["depth_fine"]) / np.abs(ref["depth_fine"])
        out["correction_%d" % corr] = {
            "depth_rel": {"median": float(np.median(rel)), "p99": float(np.quantile(rel, 0.99)),
                          "p999": float(np.quantile(rel, 0.999)), "max": float(rel.max())},
            "depth_fine_rel": {"median": float(np.median(relf)), "p99": float(np.quantile(relf, 0.99)),
                               "p999": float(np.quantile(relf, 0.999)), "max": float(relf.max())},
            "losses_rel": {k: float(abs(o[k] - ref[k]) / abs(ref[k])) for k in
                           ("child_free_loss", "child_depth_loss", "child_free_loss_fine", "child_depth_loss_fine")},
            "fwd_bwd_ms": ms}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
