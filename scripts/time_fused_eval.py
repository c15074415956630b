"""Eval-mode precision-1 MLP forward: fused single kernel (k_tc_fused_eval) vs layered row GEMMs, device time per call
from the library's own CUDA events (kernel class mlp_gemm_fwd), as TFLOP/s of the 982,528 FLOP/sample forward."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch  # noqa: E402

from pcnerf_b200 import ops  # noqa: E402
from pcnerf_b200.nof.networks import NOF_coarse  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
mc = NOF_coarse().to(dev).eval()
mc.precision = "tc"
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
out = {"peak_tflops_sustained": peak["bf16_tflops_sustained"], "peak_tflops_burst": peak["bf16_tflops"], "runs": []}
for rows in (262144, 1 << 20, 1 << 22):
    enc = (torch.randn(rows, 64, device=dev) * 0.7).half()
    enc[:, 63] = 0
    for fused in (2, 1, 0):
        ops.tc_fused_eval(fused)
        with torch.no_grad():
            for _ in range(2):
                mc.forward_encoded(enc, 1 << 22)
            torch.cuda.synchronize()
            ops.profile(True)
            n = 5
            for _ in range(n):
                mc.forward_encoded(enc, 1 << 22)
            torch.cuda.synchronize()
            prof = ops.profile_read()
            ops.profile(False)
        ms = prof["mlp_gemm_fwd"][0] / n
        small = prof["mlp_small"][0] / n
        tf = rows * 982528.0 / (ms * 1e-3) / 1e12
        out["runs"].append({"rows": rows, "engine": ("layered", "fused", "fused, CTA pairs")[fused], "gemm_ms": ms, "small_ms": small,
                            "tflops": tf, "frac_sustained": tf / peak["bf16_tflops_sustained"]})
ops.tc_fused_eval(2)
print(json.dumps(out))
