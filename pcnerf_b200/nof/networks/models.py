"""Drop-in for nof/networks/models.py: same classes, constructor arguments, parameter / buffer names and state-dict
keys (so reference checkpoints and `load_ckpt` work unchanged); `forward` runs the sm_100a kernels.

Reference facts preserved (verified against the reference by tests/golden):
  * every `nn.LeakyReLU(True)` has negative_slope == True == 1.0 -> identity (models.py:152,172,232,252);
  * the activations meant for layer2 were appended to layer1 (models.py:172) -> layer2 is Linear+BN only;
  * BatchNorm1d uses the batch statistics of the rows it is called with when the module is in train mode.
The module tree below is built the same way so that `state_dict()` keys are identical
(layer1.{0,1,3,4,6,7,9,10}.*, layer2.{0..7}.*, occ_out.0.*).
"""
import math

import torch
import torch.nn as nn

from ... import ops

_DEFAULT_PRECISION = {"value": 0}
_PREC = {"fp32": 0, "tc": 1, "bf16": 1, "f16": 1, "affine": 2, 0: 0, 1: 1, 2: 2}


def set_default_mlp_precision(p):
    """'fp32' (CUDA-core GEMM, 1e-5 parity gate) or 'tc' (tcgen05 tensor-core GEMM: fp16 operands forward, bf16
    gradients, fp32 accumulation; 1e-3 gate).  'bf16' is accepted as an alias of 'tc' (BASELINE.json's name).
    'affine': the closed form of the network as the reference builds it (identity activations): per BN batch the logit
    is exactly alpha . x + c (csrc/affine.cu); opt-in fast mode, gated at 1e-5 like 'fp32'."""
    _DEFAULT_PRECISION["value"] = _PREC[p]


class Embedding(nn.Module):
    """nof/networks/models.py:4-41: x -> (x, sin(2^k x), cos(2^k x))_{k<N_freq}."""

    def __init__(self, in_channels, N_freq, logscale=True):
        super(Embedding, self).__init__()
        self.N_freq = N_freq
        self.in_channels = in_channels
        self.funcs = [torch.sin, torch.cos]
        if logscale:
            self.freq_bands = 2 ** torch.linspace(0, N_freq - 1, N_freq)
        else:
            self.freq_bands = torch.linspace(1, 3 ** (N_freq - 1), N_freq)
        self.logscale = logscale

    def forward(self, x):
        if not (self.in_channels == 3 and self.N_freq == 10 and self.logscale):
            raise NotImplementedError("pcnerf_b200.Embedding: only in_channels=3, N_freq=10, logscale=True "
                                      "(the PC-NeRF configuration) has a kernel")
        return ops.embed(x.reshape(-1, 3), 63)


class NOF(nn.Module):
    """nof/networks/models.py:44-123 (NOF_coarse :125-203, NOF_fine :205-282, NOF_plusfine :284-359 are identical)."""

    def __init__(self, feature_size=256, in_channels_xy=63, use_skip=True):
        super(NOF, self).__init__()
        self.feature_size = feature_size
        self.in_channels_xy = in_channels_xy
        self.use_skip = use_skip
        layer1 = []
        for i in range(4):
            layer1.append(nn.Linear(in_channels_xy if i == 0 else feature_size, feature_size))
            layer1.append(nn.BatchNorm1d(num_features=feature_size))
            layer1.append(nn.LeakyReLU(True))
        layer2 = []
        for i in range(4):
            if i == 0:
                layer2.append(nn.Linear(in_channels_xy + feature_size if use_skip else feature_size, feature_size))
            else:
                layer2.append(nn.Linear(feature_size, feature_size))
            layer2.append(nn.BatchNorm1d(num_features=feature_size))
            layer1.append(nn.LeakyReLU(True))          # sic: the reference appends these to layer1
        self.layer1 = nn.Sequential(*layer1)
        self.layer2 = nn.Sequential(*layer2)
        self.occ_out = nn.Sequential(nn.Linear(feature_size, 1), nn.Sigmoid())
        self.precision = None                          # None -> module default (set_default_mlp_precision)

    # ---- kernel-facing views of the module state
    def _linears(self):
        return [self.layer1[0], self.layer1[3], self.layer1[6], self.layer1[9],
                self.layer2[0], self.layer2[2], self.layer2[4], self.layer2[6], self.occ_out[0]]

    def _bns(self):
        return [self.layer1[1], self.layer1[4], self.layer1[7], self.layer1[10],
                self.layer2[1], self.layer2[3], self.layer2[5], self.layer2[7]]

    def kernel_params(self):
        """34 tensors in the order of ops.GRAD_SIZES: (W,b) x 9 then (gamma,beta) x 8."""
        out = []
        for lin in self._linears():
            out += [lin.weight, lin.bias]
        for bn in self._bns():
            out += [bn.weight, bn.bias]
        return out

    def _buffers3(self):
        bns = self._bns()
        return ([b.running_mean for b in bns], [b.running_var for b in bns], [b.num_batches_tracked for b in bns])

    def _check_supported(self):
        if not (self.feature_size == 256 and self.in_channels_xy == 63 and self.use_skip):
            raise NotImplementedError("pcnerf_b200.NOF: kernels exist for feature_size=256, in_channels_xy=63, "
                                      "use_skip=True (the PC-NeRF configuration)")
        for bn in self._bns():
            if bn.momentum != 0.1 or bn.eps != 1e-5 or not bn.affine or not bn.track_running_stats:
                raise NotImplementedError("pcnerf_b200.NOF: BatchNorm1d must keep its default configuration")
        for m in list(self.layer1) + list(self.layer2):
            if isinstance(m, nn.LeakyReLU) and float(m.negative_slope) != 1.0:
                raise NotImplementedError("pcnerf_b200.NOF: negative_slope != 1 (the reference builds LeakyReLU(True))")

    def mlp_precision(self):
        return _DEFAULT_PRECISION["value"] if self.precision is None else _PREC[self.precision]

    # ---- closed form ("affine" mode): parameter-sized algebra in float64, batched over the BN chunks
    def _affine_coeffs(self, m, C, cnt):
        """(m (nc,64), C (nc,64,64), cnt (nc,)) float64 batch moments of the encodings -> alpha (nc,64), c (nc,) such that
        logit = alpha . x + c reproduces Linear->BN x8 ->Linear of models.py:183-203 with train-mode batch statistics
        (m is None: eval mode, running statistics).  Updates the BN running statistics in train mode."""
        lins, bns = self._linears(), self._bns()
        f64 = torch.float64
        train = m is not None
        if train:
            m63, C63 = m[:, :63], C[:, :63, :63]
            nc = m.shape[0]
        else:
            nc = 1
        A = d = None
        eps, mom = bns[0].eps, bns[0].momentum
        for l in range(8):
            W, b = lins[l].weight.to(f64), lins[l].bias.to(f64)
            if l == 0:
                A, d = W.unsqueeze(0).expand(nc, -1, -1), b.unsqueeze(0).expand(nc, -1)
            elif l == 4:
                A = W[:, :63].unsqueeze(0) + torch.matmul(W[:, 63:], Ab)
                d = torch.matmul(db, W[:, 63:].t()) + b
            else:
                A = torch.matmul(W, Ab)
                d = torch.matmul(db, W.t()) + b
            bn = bns[l]
            if train:
                mean = torch.matmul(A, m63.unsqueeze(-1)).squeeze(-1) + d
                var = (torch.matmul(A, C63) * A).sum(-1).clamp_min(0.0)
                with torch.no_grad():
                    # one running-statistics update per chunk, in chunk order (momentum recurrence in closed form)
                    k = torch.arange(nc - 1, -1, -1, device=mean.device, dtype=f64)
                    wts = mom * torch.exp(k * math.log(1.0 - mom))      # (1-mom)^k without a host scalar tensor
                    unb = var * (cnt / (cnt - 1.0)).unsqueeze(-1)
                    keep = (1.0 - mom) ** nc
                    bn.running_mean.copy_((keep * bn.running_mean.to(f64) + (wts[:, None] * mean).sum(0)).to(bn.running_mean.dtype))
                    bn.running_var.copy_((keep * bn.running_var.to(f64) + (wts[:, None] * unb).sum(0)).to(bn.running_var.dtype))
                    bn.num_batches_tracked += nc
            else:
                mean = bn.running_mean.to(f64).unsqueeze(0)
                var = bn.running_var.to(f64).unsqueeze(0)
            a = bn.weight.to(f64) / torch.sqrt(var + eps)
            s_ = bn.bias.to(f64) - mean * a
            Ab = a.unsqueeze(-1) * A
            db = a * d + s_
        wo, bo = self.occ_out[0].weight.to(f64)[0], self.occ_out[0].bias.to(f64)[0]
        alpha = torch.matmul(wo.unsqueeze(0).unsqueeze(0), Ab).squeeze(1)            # (nc,63)
        c = (db * wo).sum(-1) + bo
        alpha = torch.nn.functional.pad(alpha, (0, 1))
        return alpha, c

    def _forward_affine(self, enc, chunk):
        rows = enc.shape[0]
        if self.training:
            last = rows - (-(-rows // chunk) - 1) * chunk
            if last == 1 or chunk == 1:
                raise ValueError("Expected more than 1 value per channel when training, got input size [1, 256]")
            with torch.no_grad():
                m, C, cnt = ops.affine_moments(enc, chunk)
            alpha, c = self._affine_coeffs(m, C, cnt)
            return ops.AffineApplyFunction.apply(enc, alpha.to(torch.float32), c.to(torch.float32), chunk)
        alpha, c = self._affine_coeffs(None, None, None)
        return ops.AffineApplyFunction.apply(enc, alpha.to(torch.float32), c.to(torch.float32), rows)

    def forward_encoded(self, enc, chunk=None):
        """enc: (rows, 64) encodings padded with a zero column (fp32, or fp16 for the tensor-core path).
        One BN batch per `chunk` rows.  Returns p_occ (rows,)."""
        self._check_supported()
        prec = self.mlp_precision()
        if isinstance(enc, ops.LazyEnc):
            # (ray, depth) rows of a sampling pass (nof/render.py builds them for precision-2 models): the training pass of the
            # closed-form engine never materialises the encodings; everything else gets the tensor K2 would have written
            rows = enc.shape[0]
            chunk = rows if chunk is None else max(int(chunk), 1)
            if prec == 2 and self.training:
                last = rows - (-(-rows // chunk) - 1) * chunk
                if last == 1 or chunk == 1:
                    raise ValueError("Expected more than 1 value per channel when training, got input size [1, 256]")
                return ops.AffineRaysFunction.apply(enc.rays, enc.z, chunk, self._buffers3(), *self.kernel_params())
            if prec == 2 and not self.training and not torch.is_grad_enabled():
                # eval mode (running statistics): alpha depends on the parameters only and is cached per parameter version
                if not hasattr(self, "_eval_fold_cache"):
                    object.__setattr__(self, "_eval_fold_cache", {})
                alpha = ops.affine_eval_alpha(self.kernel_params(), self._buffers3(), self._eval_fold_cache)
                return ops.affine_apply_rays(enc.rays, enc.z, alpha)
            enc = enc.materialise()
        want = torch.float16 if prec == 1 else torch.float32
        if enc.dtype != want:
            enc = enc.to(want)
        rows = enc.shape[0]
        chunk = rows if chunk is None else int(chunk)
        if prec == 2:
            return self._forward_affine(enc.contiguous(), max(chunk, 1))
        if not hasattr(self, "_eval_fold_cache"):
            object.__setattr__(self, "_eval_fold_cache", {})       # folded eval-mode weights (ops.MLPFunction), not module state
        return ops.MLPFunction.apply(enc.contiguous(), max(chunk, 1), self.training, prec, self._buffers3(),
                                     self._eval_fold_cache,
                                     *self.kernel_params())

    def forward(self, x):
        """x: (B, 63) embedded positions -> p_occ (B, 1)   (models.py:183-203)."""
        if x.dim() != 2 or x.shape[1] != 63:
            raise ValueError("NOF.forward expects (B, 63) encodings")
        enc = torch.nn.functional.pad(x, (0, 1))
        return self.forward_encoded(enc).view(-1, 1)


class NOF_coarse(NOF):
    pass


class NOF_fine(NOF):
    pass


class NOF_plusfine(NOF):
    pass
