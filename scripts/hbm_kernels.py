"""Stand-alone launch timing of the HBM-bound kernels (K2 sample+encode, K2' resample+merge+encode, K4 compositing
forward/backward) at a batch large enough that the launch ramp does not dominate: BASELINE.json configs[3] per GPU
(32,768 rays x 128 + 256 samples) by default, `--rays N --S s --Ni ni` for other shapes.  Prints one JSON line:
algorithmic bytes / CUDA-event time per launch, as GB/s and as a fraction of MEASURED_PEAKS.json.

    python scripts/hbm_kernels.py [--rays 32768] [--S 128] [--Ni 256] [--f16 1]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pcnerf_b200 import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=32768)
    ap.add_argument("--S", type=int, default=128)
    ap.add_argument("--Ni", type=int, default=256)
    ap.add_argument("--f16", type=int, default=1)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--flush", choices=["read", "write"], default="read",
                    help="how the 126 MB L2 is emptied between launches: by READING a 256 MB buffer (clean lines) or by WRITING it "
                         "(dirty lines: their write-back to HBM then runs during, and is charged to, the timed kernel)")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    try:
        peak = float(json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except (OSError, ValueError, KeyError):
        peak = 6650.0
    n, S, Ni = a.rays, a.S, a.Ni
    F = S + Ni
    g = torch.Generator(device="cpu").manual_seed(1)
    rays = torch.zeros(n, 15)
    rays[:, 0:3] = (torch.rand(n, 3, generator=g) - 0.5) * 20.0
    d = torch.randn(n, 3, generator=g)
    rays[:, 3:6] = d / d.norm(dim=1, keepdim=True)
    rays[:, 6] = 0.5
    rays[:, 7] = 20.0 + torch.rand(n, generator=g) * 30.0
    rng = 2.5 + torch.rand(n, generator=g) * 15.0
    rays[:, 10], rays[:, 11], rays[:, 14] = rng - 0.5, rng + 0.5, rng
    rays[:, 12], rays[:, 13] = rays[:, 10], rays[:, 11]
    rays = rays.to(dev)
    esz = 128.0 if a.f16 else 256.0
    flush = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)        # > 126 MB L2, read (or rewritten) between launches
    flush32 = flush.view(torch.int32)

    def l2_flush():
        if a.flush == "write":
            flush.zero_()
        else:
            flush32.sum()            # reads 256 MB: the L2 ends up full of clean, unrelated lines

    def timed(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(a.reps):
            l2_flush()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / a.reps

    out = {"rays": n, "S": S, "Ni": Ni, "rows_f16": bool(a.f16), "peak_gbs": peak, "l2": "256 MB %s between launches" % ("read (clean lines)" if a.flush == "read" else "written (dirty lines)"),
           "kernels": {}}

    def report(name, ms, nbytes):
        gbs = nbytes / (ms * 1e-3) / 1e9
        out["kernels"][name] = {"us_per_launch": ms * 1e3, "algorithmic_mb": nbytes / 1e6, "gbs": gbs, "hbm_frac": gbs / peak}

    U = torch.rand(n, S, device=dev)
    ms = timed(lambda: ops.sample_encode_coarse(rays, S, 0, 6, 7, 10, 11, False, 1.0, U, want_enc=True, f16=bool(a.f16)))
    report("k_sample_encode_coarse", ms, n * (60.0 + S * (4.0 + 4.0 + esz)))        # + 4 B/sample of U
    z, _ = ops.sample_encode_coarse(rays, S, 0, 6, 7, 10, 11, False, 1.0, U, want_enc=False)
    p = torch.rand(n, S, device=dev) * 0.2
    w, depth, _, _, *_ = ops.composite(p, z, rays, (10, 11, 14), None, 0.0, 1e-10, ops.COMP_CHILD_LOSS)
    u = torch.rand(n, Ni, device=dev)
    ms = timed(lambda: ops.sample_encode_fine(rays, z, w, Ni, u, False, want_enc=True, f16=bool(a.f16)))
    report("k_sample_encode_fine", ms, n * (60.0 + 8.0 * S + 4.0 * Ni + F * (4.0 + esz)))
    zf, _ = ops.sample_encode_fine(rays, z, w, Ni, u, False, want_enc=False)
    pf = (torch.rand(n, F, device=dev) * 0.1).requires_grad_(True)
    for name, pp, zz, P in (("coarse", p.clone().requires_grad_(True), z, S), ("fine", pf, zf, F)):
        res = {}

        def fwd():
            res["o"] = ops.composite(pp, zz, rays, (10, 11, 14), None, 0.0, 1e-10, ops.COMP_CHILD_LOSS)

        ops.profile(True)
        ms = timed(fwd)
        # the autograd wrapper adds small torch kernels around the launch: take the library's own event time per launch
        prof = ops.profile_read()
        ops.profile(False)
        kms, kn, kb = prof["composite_fwd"]
        # (kms also holds the one-thread k_composite_losses launch that follows every forward: ~2 us)
        report("k_composite_fwd_" + name, kms / (a.reps + 2), n * (60.0 + 12.0 * P + 4.0))
        o = res["o"]
        loss = o[1].sum() + o[2] + o[3]
        ops.profile(True)
        for _ in range(a.reps):
            l2_flush()
            loss.backward(retain_graph=True)
        torch.cuda.synchronize()
        prof = ops.profile_read()
        ops.profile(False)
        kms, kn, kb = prof["composite_bwd"]
        report("k_composite_bwd_" + name, kms / a.reps, n * (60.0 + 16.0 * P))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
