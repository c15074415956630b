"""GPU: the precision-1 (TMA + tcgen05 + TMEM) path.  First the GEMM building blocks against torch.matmul on the same
16-bit operands (fp32 accumulation on both sides: agreement to fp32 summation order), then the MLP forward/backward
against the oracle, then the render path at the 1e-3 gate BASELINE.json states for the tensor-core MLP path."""
import numpy as np
import pytest
import torch

import pcnerf_oracle as orc
from conftest import golden
from gpu_util import dev, make_nets

pytestmark = pytest.mark.gpu


def _rand(shape, dt, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dt).to(dev())


@pytest.fixture(params=[0, 7], ids=["cta", "pairs"])
def row_form(request):
    """Both forms of the training-mode row GEMMs: column-split CTAs (k_tc_rowgemm) and CTA pairs (cta_group::2,
    k_tc_rowgemm2 / k_tc_wgrad2) -- bit 0 = forward, bit 1 = data gradient, bit 2 = weight gradient."""
    from pcnerf_b200 import ops
    old = ops.tc_row_pairs()
    ops.tc_row_pairs(request.param)
    yield request.param
    ops.tc_row_pairs(old)


@pytest.mark.parametrize("rows,k0,k1", [(128, 64, 0), (128, 256, 0), (1000, 256, 0), (4096 + 77, 64, 256), (40000, 256, 0)])
def test_rowgemm_forward_fp16(rows, k0, k1, row_form):
    from pcnerf_b200 import ops
    A0 = _rand((rows, k0), torch.float16, 1)
    A1 = _rand((rows, k1), torch.float16, 2) if k1 else None
    B = _rand((256, k0 + k1), torch.float16, 3, 0.1)
    bias = _rand((256,), torch.float32, 4)
    out, out2, stats = ops.tc_rowgemm(0, A0, B, A1, bias)
    A = A0 if A1 is None else torch.cat([A0, A1], 1)
    ref = A.float() @ B.float().t() + bias
    torch.cuda.synchronize()
    np.testing.assert_allclose(out.float().cpu().numpy(), ref.cpu().numpy(), rtol=2e-3, atol=2e-3)   # fp16 output rounding
    # the statistics are those of the fp16 values the next layer consumes
    o64 = out.double()
    np.testing.assert_allclose(stats[0].cpu().numpy(), o64.sum(0).cpu().numpy(), rtol=1e-5, atol=1e-3)
    np.testing.assert_allclose(stats[1].cpu().numpy(), (o64 ** 2).sum(0).cpu().numpy(), rtol=1e-5, atol=1e-3)
    assert out2 is None


@pytest.mark.parametrize("rows", [128, 300, 20000])
def test_rowgemm_dgrad_bf16_fused_bn_backward(rows, row_form):
    from pcnerf_b200 import ops
    A = _rand((rows, 256), torch.bfloat16, 5)
    B = _rand((256, 256), torch.bfloat16, 6, 0.1)
    E = _rand((rows, 256), torch.float16, 7)
    vec = _rand((4, 256), torch.float32, 8)
    out, _, stats = ops.tc_rowgemm(1, A, B, None, vec, E)
    C = A.float() @ B.float().t()
    ref = vec[0] * C - vec[1] - (E.float() - vec[3]) * vec[2]
    # bf16 output rounding; the pair form (k_tc_rowgemm2<DGRAD2>) folds vec[0] into the bf16 weight operand (here from an
    # already-bf16 B: one more 2^-9 rounding per weight) and takes the (E - mean) c2 term on the tensor core
    np.testing.assert_allclose(out.float().cpu().numpy(), ref.cpu().numpy(), rtol=1e-2, atol=5e-2 if row_form else 2e-2)
    np.testing.assert_allclose(stats[0].cpu().numpy(), out.double().sum(0).cpu().numpy(), rtol=1e-5, atol=2e-3)


@pytest.mark.parametrize("rows,ncols,xdt", [(64, 256, torch.bfloat16), (64, 256, torch.float16), (1000, 256, torch.float16),
                                            (30000, 64, torch.float16), (70000, 256, torch.float16)])
@pytest.mark.parametrize("wg_pairs", [0, 4], ids=["cta", "pairs"])
def test_wgrad_mn_major(rows, ncols, xdt, wg_pairs):
    """DH^T X with both operands MN-major.  tcgen05 kind::f16 cannot mix fp16 and bf16 operands (measured: illegal
    instruction), so an fp16 X is rewritten as bf16 tile by tile in shared memory by the kernel's idle epilogue warps:
    the reference result for fp16 X is therefore computed from bf16(X).  wg_pairs = 4: the N = 256 products on CTA pairs
    (k_tc_wgrad2: one accumulator per pair, half the split-K atomics)."""
    from pcnerf_b200 import ops
    DH = _rand((rows, 256), torch.bfloat16, 8)
    X = _rand((rows, ncols), xdt, 9)
    out = torch.zeros((256, 320), dtype=torch.float32, device=dev())
    old = ops.tc_row_pairs()
    try:
        ops.tc_row_pairs((old & 3) | wg_pairs)
        ops.tc_wgrad(DH, X, ncols, out, 64 if ncols == 256 else 0)
        torch.cuda.synchronize()
    finally:
        ops.tc_row_pairs(old)
    assert ops.lib().pcnerf_tc_last_fault() == 0
    ref = DH.double().t() @ X.to(torch.bfloat16).double()
    off = 64 if ncols == 256 else 0
    got = out[:, off:off + ncols].double()
    scale = float(ref.abs().max())
    np.testing.assert_allclose(got.cpu().numpy(), ref.cpu().numpy(), rtol=1e-4, atol=1e-5 * scale)
    rest = out.clone()
    rest[:, off:off + ncols] = 0
    assert float(rest.abs().max()) == 0.0


DEFAULT_FUSED = 2


def _enc(rows, seed):
    gen = torch.Generator().manual_seed(seed)
    x = (torch.rand(rows, 3, generator=gen) - 0.5) * 60.0
    return orc.embedding(x)


@pytest.mark.parametrize("rows,chunk", [(4096, 4096), (5000, 2048), (130, 130)])
def test_mlp_forward_tc(rows, chunk, row_form):
    enc = _enc(rows, rows)
    sd = orc.init_state_dict(42)
    p_ref = torch.cat([orc.nof_forward(sd, enc[i:i + chunk], True) for i in range(0, rows, chunk)]).reshape(-1)
    mc, _, _ = make_nets(42, 43, True, "tc")
    p = mc.forward_encoded(torch.nn.functional.pad(enc, (0, 1)).to(dev()), chunk)
    err = (p.detach().cpu() - p_ref).abs() / p_ref
    assert float(err.max()) < 6e-3 and float(err.mean()) < 1e-3, (float(err.max()), float(err.mean()))
    got = mc.state_dict()
    for k in ("layer1.1.running_mean", "layer1.1.running_var", "layer2.7.running_mean", "layer2.7.running_var"):
        np.testing.assert_allclose(got[k].cpu().numpy(), sd[k].numpy(), rtol=5e-3, atol=5e-3, err_msg=k)


@pytest.mark.parametrize("rows,chunk", [(4096, 4096), (3000, 1024)])
def test_mlp_backward_tc(rows, chunk, row_form):
    enc = _enc(rows, rows + 1)
    gen = torch.Generator().manual_seed(rows)
    gp = torch.randn(rows, generator=gen)
    sd = orc.init_state_dict(42)
    for k in orc.param_names():
        sd[k].requires_grad_(True)
    p_ref = torch.cat([orc.nof_forward(sd, enc[i:i + chunk], True) for i in range(0, rows, chunk)]).reshape(-1)
    (p_ref * gp).sum().backward()
    mc, _, _ = make_nets(42, 43, True, "tc")
    p = mc.forward_encoded(torch.nn.functional.pad(enc, (0, 1)).to(dev()), chunk)
    (p * gp.to(dev())).sum().backward()
    scale = max(float(sd[k].grad.abs().max()) for k in orc.param_names())
    for k, prm in mc.named_parameters():
        ref = sd[k].grad.numpy()
        got = prm.grad.cpu().numpy()
        # 16-bit operands: compare each tensor on its own scale (bf16 gradients: ~1 % of the tensor's max)
        atol = 2e-2 * np.abs(ref).max() + 1e-12
        if k.endswith(".bias") and k.split(".")[0] in ("layer1", "layer2") and k != "layer2.7.bias":
            atol = 1e-3 * scale          # exactly zero in exact arithmetic: rounding noise of the 16-bit path
        assert np.abs(got - ref).max() <= atol, (k, np.abs(got - ref).max(), atol)
        if ref.size > 256:
            cos = float((got * ref).sum() / (np.linalg.norm(got) * np.linalg.norm(ref) + 1e-30))
            assert cos > 0.999, (k, cos)


@pytest.mark.parametrize("name", ["train_seg", "train_plain"])
def test_train_coarse_pass_tc_gate_1e3(name):
    """BASELINE.json: rendered depth and losses within 1e-3 relative for the tensor-core MLP path (coarse pass; the
    fine pass inherits the resampling conditioning documented in tests/test_gpu_render.py)."""
    from pcnerf_b200.nof import render
    g = golden(name)
    rays = torch.from_numpy(g["rays"]).to(dev())
    mc, mf, emb = make_nets(42, 43, True, "tc")
    res = render.render_rays_train(mc, mf, emb, rays, N_samples=int(g["S"]), N_importance=int(g["Ni"]), perturb=0,
                                   noise_std=0, chunk=int(g["chunk"]), issegmentated=int(g["issegmentated"]),
                                   childnerf_ratio=float(g["ratio"]), use_child_nerf_divide=0,
                                   use_child_nerf_loss=int(g["use_child"]))
    np.testing.assert_allclose(res["depth"].detach().cpu().numpy(), g["out_depth"], rtol=1e-3, atol=1e-6)
    if int(g["use_child"]):
        np.testing.assert_allclose(float(res["child_free_loss"]), g["out_child_free_loss"], rtol=1e-3)
        np.testing.assert_allclose(float(res["child_depth_loss"]), g["out_child_depth_loss"], rtol=1e-3)
    np.testing.assert_allclose(res["depth_fine"].detach().cpu().numpy(), g["out_depth_fine"], rtol=5e-3, atol=1e-6)


def test_mlp_eval_tc_chunking_is_invisible():
    """Eval-mode BN uses running statistics, so `chunk` must not change a single bit: the tensor-core path derives the
    folded weights once per pass (first chunk), later chunks reuse them and no chunk computes batch statistics.  Checked
    with non-trivial running statistics (one training pass first) against the fp32 oracle at the 1e-3 gate."""
    from pcnerf_b200 import ops
    rows = 3000
    enc = _enc(rows, 7)
    mc, _, _ = make_nets(42, 43, True, "tc")
    encd = torch.nn.functional.pad(enc, (0, 1)).to(dev())
    mc.forward_encoded(encd, 1024)                       # train-mode pass: running statistics move away from (0, 1)
    mc.eval()
    sd = {k: v.detach().cpu().clone() for k, v in mc.state_dict().items()}
    p_ref = orc.nof_forward(sd, enc, False).reshape(-1)
    floor = ops.EVAL_CHUNK_FLOOR
    try:
        ops.EVAL_CHUNK_FLOOR = 0
        with torch.no_grad():
            p_chunks = mc.forward_encoded(encd, 1024).cpu()          # 3 chunks: 1024, 1024, 952 rows
            p_odd = mc.forward_encoded(encd, 777).cpu()
        ops.EVAL_CHUNK_FLOOR = floor
        with torch.no_grad():
            p_one = mc.forward_encoded(encd, 1024).cpu()             # one launch
    finally:
        ops.EVAL_CHUNK_FLOOR = floor
    assert torch.equal(p_chunks, p_one) and torch.equal(p_odd, p_one)
    err = (p_one - p_ref).abs() / p_ref
    assert float(err.max()) < 6e-3 and float(err.mean()) < 1e-3, (float(err.max()), float(err.mean()))
    for k in ("layer1.1.running_mean", "layer2.7.running_var"):    # eval mode leaves the running statistics alone
        assert torch.equal(mc.state_dict()[k].cpu(), sd[k])


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("rows", [127, 130, 3000, 100003])
def test_mlp_eval_fused_matches_layered(rows, mode):
    """k_tc_fused_eval (all layers in one kernel, activations in shared memory / TMEM) against the layered eval path (same
    fp16 operands, same folded weights: the hidden activations are bit-identical, only the 256-term output dot is summed in
    a different order) and against the fp32 oracle at the tensor-core gate.  Row counts: a lone partial tile (the pair's
    second tile is entirely out of bounds), a partial second tile, many pairs, and more pairs than CTAs (barrier phases
    wrap).  mode 1 = one CTA per unit of work, mode 2 = CTA pairs (cta_group::2 MMAs, M = 256)."""
    from pcnerf_b200 import ops
    enc = _enc(rows, rows + 3)
    mc, _, _ = make_nets(42, 43, True, "tc")
    encd = torch.nn.functional.pad(enc, (0, 1)).to(dev())
    mc.forward_encoded(encd[:4096], 4096)                # one train-mode pass: non-trivial running statistics
    mc.eval()
    sd = {k: v.detach().cpu().clone() for k, v in mc.state_dict().items()}
    try:
        ops.tc_fused_eval(0)
        ops.tc_weight_correction(0)      # the fused kernel streams the plain fp16(W') operands: compare like with like
        with torch.no_grad():
            p_lay = mc.forward_encoded(encd, 1 << 20).cpu()
        ops.tc_fused_eval(mode)
        with torch.no_grad():
            p_fus = mc.forward_encoded(encd, 1 << 20).cpu()
    finally:
        ops.tc_fused_eval(DEFAULT_FUSED)
        ops.tc_weight_correction(1)
    assert ops.lib().pcnerf_tc_last_fault() == 0
    np.testing.assert_allclose(p_fus.numpy(), p_lay.numpy(), rtol=2e-5, atol=1e-7)
    n_ref = min(rows, 4096)
    p_ref = orc.nof_forward(sd, enc[:n_ref], False).reshape(-1)
    err = (p_fus[:n_ref] - p_ref).abs() / p_ref
    assert float(err.max()) < 6e-3 and float(err.mean()) < 1e-3, (float(err.max()), float(err.mean()))


def test_two_chunks_in_flight_is_bit_identical():
    """Training-mode tensor-core MLP with two BN chunks in flight on internal streams (ops.TC_LANES = 2) against one chunk at
    a time: outputs and running statistics must be bit-identical (the running-statistics updates are event-ordered in chunk
    order across the two lanes, like the += into the gradient buffers).  The gradients themselves are not bit-reproducible
    from run to run even with one lane (split-K fp32 atomics in the weight-gradient GEMM, then bf16 rounding downstream):
    they are compared at the run-to-run level, on each tensor's scale."""
    from pcnerf_b200 import ops
    rows, chunk = 5000, 1024                                   # 5 chunks, the last one partial
    enc = torch.nn.functional.pad(_enc(rows, 99), (0, 1)).to(dev())
    gen = torch.Generator().manual_seed(1)
    gp = torch.randn(rows, generator=gen).to(dev())
    res = {}
    lanes0 = ops.TC_LANES
    try:
        for lanes in (1, 2):
            ops.TC_LANES = lanes
            mc, _, _ = make_nets(42, 43, True, "tc")
            p = mc.forward_encoded(enc, chunk)
            (p * gp).sum().backward()
            torch.cuda.synchronize()
            res[lanes] = (p.detach().clone(), {k: v.grad.clone() for k, v in mc.named_parameters()},
                          {k: v.clone() for k, v in mc.state_dict().items() if "running" in k or "num_batches" in k})
    finally:
        ops.TC_LANES = lanes0
    assert torch.equal(res[1][0], res[2][0])
    for k in res[1][2]:
        assert torch.equal(res[1][2][k], res[2][2][k]), k
    for k in res[1][1]:
        a, b = res[1][1][k], res[2][1][k]
        if k.endswith(".bias") and k.split(".")[0] in ("layer1", "layer2") and k != "layer2.7.bias":
            continue                                       # exactly zero in exact arithmetic: rounding noise either way
        assert float((a - b).abs().max()) <= 1e-2 * float(a.abs().max()) + 1e-12, k


def test_eval_fold_cache_follows_parameter_updates():
    """The folded eval-mode weights are cached per model and re-derived only when a parameter or a running statistic was
    written to: two models alternating (coarse / fine) must not see each other's weights or epilogue constants, and an
    in-place parameter update must invalidate the cache."""
    rows = 2000
    enc = torch.nn.functional.pad(_enc(rows, 5), (0, 1)).to(dev())
    ma, mb, _ = make_nets(42, 43, False, "tc")

    def ref(m):
        sd = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
        return orc.nof_forward(sd, enc.cpu()[:, :63].float(), False).reshape(-1)

    def close(p, r):
        err = (p.cpu() - r).abs() / r
        return float(err.max()) < 6e-3 and float(err.mean()) < 1e-3

    with torch.no_grad():
        pa1 = mb.forward_encoded(enc, rows) * 0 + ma.forward_encoded(enc, rows)      # b then a: a must not inherit b's constants
        pb1 = mb.forward_encoded(enc, rows)
        pa2 = ma.forward_encoded(enc, rows)                                          # cached path (prepared = 2)
        assert torch.equal(pa1, pa2) and close(pa1, ref(ma)) and close(pb1, ref(mb))
        assert not torch.equal(pa1, pb1)
        ma.occ_out[0].bias.add_(0.5)                                                 # in-place update -> new version
        ma.layer1[0].weight.mul_(1.1)
        pa3 = ma.forward_encoded(enc, rows)
        assert not torch.equal(pa3, pa2) and close(pa3, ref(ma))
        ma.layer1[1].running_mean.add_(0.05)                                         # running statistics count too
        pa4 = ma.forward_encoded(enc, rows)
        assert not torch.equal(pa4, pa3) and close(pa4, ref(ma))


def test_eval_fold_cache_sees_library_side_writes():
    """FlatAdam's kernel and the running-statistics update of a training-mode forward write through raw
    pointers, which never bump torch's tensor versions -- the cached folded eval-mode weights must still be re-derived
    (train -> validate -> train -> validate)."""
    from pcnerf_b200.graphed import GraphedStep
    from pcnerf_b200.optim import FlatAdam
    rows = 3000
    enc = torch.nn.functional.pad(_enc(rows, 6), (0, 1)).to(dev())
    m, _, _ = make_nets(42, 43, False, "tc")
    opt = FlatAdam(list(m.parameters()), lr=1e-3, eps=1e-8, weight_decay=1e-3)      # (one step moves every weight by ~lr)

    def ref():
        sd = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
        return orc.nof_forward(sd, enc.cpu()[:, :63].float(), False).reshape(-1)

    def close(p, r):
        err = (p.cpu() - r).abs() / r
        return float(err.max()) < 6e-3 and float(err.mean()) < 1e-3

    def evaluate():
        m.eval()
        with torch.no_grad():
            return m.forward_encoded(enc, rows).clone()

    p0 = evaluate()
    assert torch.equal(p0, evaluate()) and close(p0, ref())          # second call: cached fold
    # (1) optimizer step by the library's Adam kernel
    opt.bucket.flat.normal_(generator=torch.Generator(device=dev()).manual_seed(3))
    opt.step()
    p1 = evaluate()
    assert not torch.equal(p1, p0) and close(p1, ref())
    # (2) a training-mode forward moves the running statistics
    m.train()
    m.forward_encoded(enc, rows)
    p2 = evaluate()
    assert not torch.equal(p2, p1) and close(p2, ref())
    # (3) the same two writes replayed from a CUDA graph (no Python runs during a replay)
    m.train()

    def step():
        opt.bucket.zero()
        m.forward_encoded(enc, rows).sum().backward()
        opt.step()

    g = GraphedStep(step, warmup=1)
    p3 = evaluate()
    g()
    p4 = evaluate()
    assert not torch.equal(p4, p3) and close(p4, ref())


def test_tc_training_tail_chunk_of_one_row_raises_before_any_launch():
    """the chunked training pass validates every chunk like pcnerf_mlp_forward does -- a tail chunk of
    exactly one row raises torch's ValueError and nothing has been launched (running statistics untouched)."""
    chunk = 512
    enc = torch.nn.functional.pad(_enc(chunk + 1, 7), (0, 1)).to(dev())
    m, _, _ = make_nets(42, 43, True, "tc")
    before = {k: v.clone() for k, v in m.state_dict().items() if "running" in k or "num_batches" in k}
    with pytest.raises(ValueError, match="Expected more than 1 value per channel"):
        m.forward_encoded(enc, chunk)
    torch.cuda.synchronize()
    for k, v in m.state_dict().items():
        if k in before:
            assert torch.equal(v, before[k]), k
    p = m.forward_encoded(enc[:chunk].contiguous(), chunk)
    assert bool(torch.isfinite(p).all())


def test_flat_adam_state_dict_roundtrip_and_grad_realias():
    """the Adam moments / step count survive state_dict() -> load_state_dict(), and a parameter whose
    .grad was detached (zero_grad(set_to_none=True)) is re-aliased instead of silently skipped."""
    from pcnerf_b200.optim import FlatAdam
    torch.manual_seed(0)
    ma, _, _ = make_nets(42, 43, True, "tc")
    mb, _, _ = make_nets(42, 43, True, "tc")
    oa = FlatAdam(list(ma.parameters()), lr=1e-3, eps=1e-8, weight_decay=1e-3)
    ob = FlatAdam(list(mb.parameters()), lr=1e-3, eps=1e-8, weight_decay=1e-3)
    gen = torch.Generator(device=dev()).manual_seed(11)
    for _ in range(3):
        oa.bucket.flat.normal_(generator=gen)
        oa.step()
    ob.load_state_dict(oa.state_dict())
    with torch.no_grad():
        ob.flat.copy_(oa.flat)
    g = torch.randn(oa.flat.numel(), device=dev(), generator=gen)
    oa.bucket.flat.copy_(g)
    oa.step()
    # b: gradients arrive in fresh tensors (set_to_none semantics)
    o = 0
    for p in ob.bucket.params:
        p.grad = g[o:o + p.numel()].view_as(p).clone()
        o += p.numel()
    ob.step()
    torch.cuda.synchronize()
    assert torch.equal(oa.flat, ob.flat) and torch.equal(oa.exp_avg, ob.exp_avg) and int(ob.step_dev.item()) == 4


@pytest.mark.parametrize("training", [True, False], ids=["train", "eval_layered"])
def test_weight_correction_removes_the_weight_rounding_term(training):
    """k_tc_fold's correction block (C = (W' - fp16(W')) T_l applied to the encoding) against the plain fp16(W') operands:
    the mean relative error of p against the fp32 oracle must drop (what remains is the fp16 rounding of the activations
    and of the encoding), in training mode (batch statistics) and in the layered eval path (running statistics)."""
    from pcnerf_b200 import ops
    rows = 8192
    enc = _enc(rows, 21)
    encd = torch.nn.functional.pad(enc, (0, 1)).to(dev())
    errs = {}
    try:
        ops.tc_fused_eval(0)
        for corr in (0, 1):
            ops.tc_weight_correction(corr)
            mc, _, _ = make_nets(42, 43, training, "tc")
            sd = {k: v.detach().cpu().clone() for k, v in mc.state_dict().items()}
            with torch.no_grad():
                p = mc.forward_encoded(encd, rows).cpu()
            ref = orc.nof_forward(sd, enc, training).reshape(-1)
            rel = ((p - ref).abs() / ref).double()
            errs[corr] = (float(rel.mean()), float(rel.max()))
            if training:                                        # running statistics updated exactly once, from the same batch
                np.testing.assert_allclose(mc.state_dict()["layer2.7.running_var"].cpu().numpy(),
                                           sd["layer2.7.running_var"].numpy(), rtol=2e-3)
    finally:
        ops.tc_fused_eval(DEFAULT_FUSED)
        ops.tc_weight_correction(1)
    assert errs[1][0] < 0.95 * errs[0][0], errs          # measured: 2.7e-4 -> 2.3e-4 (train), the rest is activation rounding
    assert errs[1][1] < 6e-3, errs
