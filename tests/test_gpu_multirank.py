"""GPU (needs >= 2 devices; skipped on a one-GPU box): ranks end a data-parallel step with identical weights, equal to the
single-process step on the whole batch -- the multi-GPU semantics the reference does not define (train_kitti.py:283-292 is
single-device) and pcnerf_b200.parallel does.  Runs tests/multirank_worker.py under torch.distributed.run with NCCL."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("precision", ["fp32", "tc"])
def test_two_ranks_identical_weights_and_global_batch_gradient(precision):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tests", "multirank_worker.py"), precision]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0 and lines, (r.returncode, r.stdout[-2000:], r.stderr[-2000:])
    out = json.loads(lines[-1])
    assert out["ok"] and out["rank_identical_weights"], out
