"""GPU: FlatAdam (one kernel over one flat buffer) against torch.optim.Adam as the reference configures
it (nof/nof_utils.py:162-173), under MultiStepLR (train_kitti.py:113) and replayed from a CUDA graph; LossHistory."""
import numpy as np
import pytest
import torch

from gpu_util import dev, make_nets

pytestmark = pytest.mark.gpu


def _pair():
    ma, mb, _ = make_nets(42, 43, True, None)
    mc, md, _ = make_nets(42, 43, True, None)
    return list(ma.parameters()) + list(mb.parameters()), list(mc.parameters()) + list(md.parameters())


def test_flat_adam_matches_torch_adam_with_scheduler():
    from pcnerf_b200.optim import FlatAdam
    pa, pb = _pair()
    ref = torch.optim.Adam(pa, lr=5e-4, eps=1e-8, weight_decay=1e-3)
    opt = FlatAdam(pb, lr=5e-4, eps=1e-8, weight_decay=1e-3)
    assert sum(p.numel() for p in pb) == opt.flat.numel() == 2 * 497409
    assert all(p.data_ptr() >= opt.flat.data_ptr() for p in pb) and pb[0].shape == pa[0].shape
    s_ref = torch.optim.lr_scheduler.MultiStepLR(ref, milestones=[3, 6], gamma=0.2)
    s_opt = torch.optim.lr_scheduler.MultiStepLR(opt, milestones=[3, 6], gamma=0.2)
    gen = torch.Generator(device="cuda").manual_seed(0)
    for it in range(8):
        opt.zero_grad()
        for a, b in zip(pa, pb):
            g = torch.randn(a.shape, device=dev(), generator=gen) * (10.0 ** (it % 3 - 1))
            a.grad = g.clone()
            b.grad.copy_(g)                                  # .grad aliases the flat bucket
        ref.step()
        opt.step()
        s_ref.step()
        s_opt.step()
    for a, b in zip(pa, pb):
        np.testing.assert_allclose(b.detach().cpu().numpy(), a.detach().cpu().numpy(), rtol=2e-6, atol=2e-7)
    assert int(opt.step_dev.item()) == 8 and abs(float(opt.lr_dev.item()) - 5e-4 * 0.04) < 1e-12


def test_flat_adam_graph_replay_and_loss_history(tmp_path):
    from pcnerf_b200.optim import FlatAdam, LossHistory
    pa, pb = _pair()
    ref = torch.optim.Adam(pa, lr=1e-3, eps=1e-8, weight_decay=1e-3)
    opt = FlatAdam(pb, lr=1e-3, eps=1e-8, weight_decay=1e-3)
    hist = LossHistory(capacity=4, device=dev())
    g = torch.randn(opt.flat.numel(), device=dev())
    o = 0
    for a in pa:
        a.grad = g[o:o + a.numel()].view_as(a).clone()
        o += a.numel()
    terms = [torch.full((1,), float(i), device=dev()) for i in range(7)]

    def body():
        opt.bucket.flat.copy_(g)
        opt.step()
        hist.append(terms)

    body()
    ref.step()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        with torch.cuda.graph(graph, stream=s):
            body()
    for _ in range(5):
        graph.replay()
        ref.step()
    torch.cuda.synchronize()
    assert int(opt.step_dev.item()) == 6
    for a, b in zip(pa, pb):
        np.testing.assert_allclose(b.detach().cpu().numpy(), a.detach().cpu().numpy(), rtol=2e-6, atol=2e-7)
    h = hist.to_numpy()
    assert h.shape == (4, 7) and np.array_equal(h[0], np.arange(7, dtype=np.float32))       # ring of 4, 6 appends
    paths = [str(tmp_path / ("h%d.npy" % i)) for i in range(7)]
    hist.save(paths)
    assert np.array_equal(np.load(paths[3]), np.full(4, 3.0, dtype=np.float32))
