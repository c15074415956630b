import numpy as np
import torch

import pcnerf_oracle as orc

STRIDE = 17
BIG = ("layer1.3.weight", "layer1.6.weight", "layer1.9.weight", "layer2.0.weight", "layer2.2.weight",
       "layer2.4.weight", "layer2.6.weight")


def dev():
    return torch.device("cuda:0")


def make_nets(seed_c=42, seed_f=43, train=True, precision=None):
    from pcnerf_b200.nof.networks import NOF_coarse, NOF_fine, Embedding
    mc, mf = NOF_coarse(), NOF_fine()
    mc.load_state_dict(orc.init_state_dict(seed_c))
    mf.load_state_dict(orc.init_state_dict(seed_f))
    mc.to(dev()).train(train)
    mf.to(dev()).train(train)
    mc.precision = precision
    mf.precision = precision
    return mc, mf, Embedding(3, 10)


def grads_compressed(model):
    out = {}
    for k, p in model.named_parameters():
        g = p.grad.detach().cpu().numpy()
        out[k] = g.reshape(-1)[::STRIDE] if k in BIG else g
    return out


def assert_grads_match(tag, model, g, rtol, noise_rel=1e-5):
    """Compare parameter gradients with a golden fixture (see tests/test_oracle_golden.py for the noise class)."""
    scale = max(np.abs(g[kk]).max() for kk in g.files if kk.startswith("grad_%s_" % tag))
    for k, gr in grads_compressed(model).items():
        ref = g["grad_%s_%s" % (tag, k)]
        atol = rtol * np.abs(ref).max() + 1e-12
        if k.endswith(".bias") and k.split(".")[0] in ("layer1", "layer2") and k != "layer2.7.bias":
            atol = noise_rel * scale
        np.testing.assert_allclose(gr, ref, rtol=rtol, atol=atol, err_msg="%s %s" % (tag, k))


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / (np.abs(b) + 1e-12)))
