"""Worker of tests/test_gpu_multirank.py (one process per GPU, launched by torch.distributed.run, NCCL).

Every rank renders ITS shard of a ray batch (per-rank, per-chunk BatchNorm statistics: the DDP semantics of
pcnerf_b200.parallel), scales the child depth loss by 1/world (nof/render.py:155 carries 1/N^2), back-propagates into the
flat gradient bucket and takes one FlatAdam step (NCCL all-reduce of the bucket + Adam).  Checks:
  (1) after the step every rank holds bit-identical weights;
  (2) the all-reduced, averaged gradient equals the gradient of ONE process rendering the whole batch (rank 0 recomputes it
      on its own GPU: the shards hold whole BatchNorm chunks, so the global batch forms the same chunks);
  (3) the updated weights equal the single-process step's.
Prints one JSON line on rank 0; exit code 0 = all checks passed."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    precision = sys.argv[1] if len(sys.argv) > 1 else "fp32"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import pcnerf_oracle as orc
    from pcnerf_b200 import parallel, synth
    from pcnerf_b200.nof import render
    from pcnerf_b200.nof.networks import Embedding, NOF_coarse, NOF_fine
    from pcnerf_b200.optim import FlatAdam

    n_l, S, NI, chunk = 1024, 64, 128, 8192            # per-rank shard: 8 coarse / 24 fine whole chunks
    rays_all = torch.from_numpy(synth.synth_train_rays(77, n_l * world, K=8)).to(dev)
    lam = (1.0, 1e6, 1e5)
    sl1 = torch.nn.SmoothL1Loss(reduction="mean")

    def nets():
        mc, mf = NOF_coarse(), NOF_fine()
        mc.load_state_dict(orc.init_state_dict(42))
        mf.load_state_dict(orc.init_state_dict(43))
        mc.to(dev).train()
        mf.to(dev).train()
        mc.precision = mf.precision = precision
        return mc, mf

    def loss_of(mc, mf, rays, dscale):
        res = render.render_rays_train(mc, mf, Embedding(3, 10), rays, N_samples=S, N_importance=NI, perturb=0, noise_std=0,
                                       chunk=chunk, issegmentated=1, childnerf_ratio=0.1, use_child_nerf_divide=0,
                                       use_child_nerf_loss=1)
        gt = rays[:, 14]
        return 0.1 * lam[0] * sl1(10 * res["depth"], 10 * gt) + 0.1 * lam[0] * sl1(10 * res["depth_fine"], 10 * gt) \
            + lam[1] * (res["child_free_loss_fine"] + res["child_free_loss"]) \
            + lam[2] * dscale * (res["child_depth_loss_fine"] + res["child_depth_loss"])

    # ---- data-parallel step
    mc, mf = nets()
    opt = FlatAdam(list(mc.parameters()) + list(mf.parameters()), lr=5e-4, eps=1e-8, weight_decay=1e-3)
    opt.zero_grad()
    a, b = parallel.shard_rows(n_l * world, world, rank)
    loss = loss_of(mc, mf, rays_all[a:b], parallel.depth_loss_scale())
    loss.backward()
    dist.all_reduce(opt.bucket.flat, op=dist.ReduceOp.SUM)
    grad_dp = (opt.bucket.flat / world).clone()                   # what the fused step applies (1/world folded in)
    # ---- the same step through FlatAdam.step (NCCL all-reduce of the flat bucket + Adam), from fresh nets
    mc, mf = nets()
    opt = FlatAdam(list(mc.parameters()) + list(mf.parameters()), lr=5e-4, eps=1e-8, weight_decay=1e-3)
    opt.zero_grad()
    loss = loss_of(mc, mf, rays_all[a:b], parallel.depth_loss_scale())
    loss.backward()
    opt.step()                                                    # NCCL all-reduce + Adam, as bench.py / fit() run it
    torch.cuda.synchronize()
    w_dp = opt.flat.clone()
    gathered = [torch.empty_like(w_dp) for _ in range(world)]
    dist.all_gather(gathered, w_dp)
    same = all(torch.equal(g, gathered[0]) for g in gathered)

    out = {"precision": precision, "world": world, "rank_identical_weights": bool(same)}
    ok = same
    if rank == 0:
        # ---- the same batch in ONE process
        mc1, mf1 = nets()
        opt1 = FlatAdam(list(mc1.parameters()) + list(mf1.parameters()), lr=5e-4, eps=1e-8, weight_decay=1e-3)
        opt1.zero_grad()
        loss1 = loss_of(mc1, mf1, rays_all, 1.0)
        loss1.backward()
        g1 = opt1.bucket.flat.clone()
        opt1.step(allreduce=False)
        torch.cuda.synchronize()
        scale = float(g1.abs().max())
        gerr = float((grad_dp - g1).abs().max()) / scale
        # per-tensor check on each tensor's own scale
        worst, o = 0.0, 0
        for p in opt1.bucket.params:
            s_ = slice(o, o + p.numel())
            t = float(g1[s_].abs().max())
            if t > 1e-6 * scale:
                worst = max(worst, float((grad_dp[s_] - g1[s_]).abs().max()) / t)
            o += p.numel()
        # the first Adam step moves every weight by ~lr * sign(g): where the exact gradient is zero (the Linear biases in front
        # of a train-mode BatchNorm) the sign is rounding noise, so the updated weights are compared where the gradient is
        # resolved (|g| > 1e-3 of the largest entry) -- everywhere for the fp32 engine's own check below
        sel = g1.abs() > 1e-3 * scale
        werr = float((w_dp - opt1.flat)[sel].abs().max())
        werr_all = float((w_dp - opt1.flat).abs().max())
        tol_g = 2e-4 if precision == "fp32" else 5e-2
        out.update({"grad_max_err_over_global_max": gerr, "grad_worst_tensor_rel_err": worst, "weights_max_abs_diff": werr,
                    "weights_max_abs_diff_incl_zero_gradient_entries": werr_all, "resolved_entries": int(sel.sum()),
                    "loss_global": float(loss1), "tolerance_grad": tol_g})
        # Adam normalises the step: |dw| <= lr per element, so weights agree to a fraction of lr
        ok = ok and gerr < tol_g and werr < (5e-5 if precision == "fp32" else 2e-4)
        if precision == "fp32":
            ok = ok and werr_all < 5e-5
        out["ok"] = bool(ok)
        print(json.dumps(out), flush=True)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
