"""Drop-in `spatial_vae.models` for NVIDIA B200.

Same classes, constructor arguments, attribute names and state_dict keys as the reference
module (reference spatial_vae/models.py:13-172), so whole-module `.sav` pickles and state
dicts interchange.  The submodules only HOLD the parameters; `forward` hands them to the
sm_100a kernels in libsvae_b200.so through `spatial_vae.functional` (no eager arithmetic, no CPU
fallback).  The options resid / expand_coords / bilinear run through option_kernels.cu (GPU parity tests:
tests/test_gpu_zz_options.py).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functional as SF


def _stack(first_in, width, depth, act, resid):
    """[Linear(first_in,width)?, act, (Linear(width,width), act) * (depth-1)] with the reference's
    Sequential index layout (reference models.py:31-39, 77-83).  first_in=None omits the leading
    Linear (the decoder's first layer lives in coord_linear/latent_linear).  Parameters are created
    in the reference's order so the same torch.manual_seed gives the same initialisation."""
    mods = ([nn.Linear(first_in, width)] if first_in is not None else []) + [act()]
    for _ in range(depth - 1):
        if resid:
            mods.append(ResidLinear(width, width, activation=act))
        else:
            mods += [nn.Linear(width, width), act()]
    return mods


class ResidLinear(nn.Module):
    """act(linear(x) + x) (reference models.py:13-21).  Inside InferenceNetwork / SpatialGenerator it is evaluated as
    part of the network's fused forward (the skip connection rides in the GEMM epilogue); called on its own it runs
    the same fp32 GEMM-with-addend through svae_resid_linear_forward / _backward (autograd supported)."""

    def __init__(self, n_in, n_out, activation=nn.Tanh):
        super().__init__()
        self.linear = nn.Linear(n_in, n_out)
        self.act = activation()

    def forward(self, x):
        return SF.resid_linear(x, self.linear.weight, self.linear.bias, SF.activation_code(type(self.act)))


class _Derived:
    """What the kernels need to know beyond the parameters is READ OFF the module structure instead of being stored,
    so that whole-module pickles written by the reference (whose objects only carry the reference's own attributes,
    models.py:28-29,62-68) evaluate here unchanged, and pickles written here carry nothing the reference lacks."""
    _ACT_INDEX = 0
    precision = None          # None = spatial_vae.functional.default_precision(); may be set per instance

    @property
    def activation_code(self):
        return SF.activation_code(type(self.layers[self._ACT_INDEX]))

    @property
    def resid(self):
        return any(isinstance(m, ResidLinear) for m in self.layers)


class InferenceNetwork(_Derived, nn.Module):
    """Encoder MLP: image (B, n) -> (z_mu, z_logstd), each (B, latent_dim) (reference models.py:24-54)."""
    _ACT_INDEX = 1            # layers = [Linear, act, ...]

    def __init__(self, n, latent_dim, hidden_dim, num_layers=1, activation=nn.Tanh, resid=False):
        super().__init__()
        self.latent_dim = latent_dim
        self.n = n
        self.layers = nn.Sequential(*_stack(n, hidden_dim, num_layers, activation, resid),
                                    nn.Linear(hidden_dim, 2 * latent_dim))
        SF.activation_code(activation)        # unsupported activations fail at construction
        print(self)

    def forward(self, x):
        out = SF.encoder_forward(self, x)
        return out[:, :self.latent_dim], out[:, self.latent_dim:]


class SpatialGenerator(_Derived, nn.Module):
    """Coordinate-conditioned decoder, evaluated once per pixel (reference models.py:57-132):
    y[b,p,:] = sigmoid(W_o h_{L-1} + b_o),  h_0 = act(W_c x[b,p] + b_c + W_z z[b]),
    h_l = act(W_l h_{l-1} + b_l)."""

    def __init__(self, latent_dim, hidden_dim, n_out=1, num_layers=1, activation=nn.Tanh,
                 softplus=False, resid=False, expand_coords=False, bilinear=False):
        super().__init__()
        self.softplus = softplus
        self.expand_coords = expand_coords
        self.latent_dim = latent_dim
        coord_features = 5 if expand_coords else 2
        self.coord_linear = nn.Linear(coord_features, hidden_dim)
        if latent_dim > 0:
            self.latent_linear = nn.Linear(latent_dim, hidden_dim, bias=False)
            if bilinear:
                self.bilinear = nn.Bilinear(coord_features, latent_dim, hidden_dim, bias=False)
        body = _stack(None, hidden_dim, num_layers, activation, resid)   # starts with the activation
        self.layers = nn.Sequential(*body, nn.Linear(hidden_dim, n_out), nn.Sigmoid())
        SF.activation_code(activation)        # unsupported activations fail at construction
        print(self)

    def forward(self, x, z):
        if x.dim() < 3:
            x = x.unsqueeze(0)
        if z is not None and z.dim() < 2:
            z = z.unsqueeze(0)
        if z is not None and z.shape[0] != x.shape[0]:
            z = z.expand(x.shape[0], -1)
        return SF.decoder_forward(self, x, z if hasattr(self, "latent_linear") else None)


class VanillaGenerator(nn.Module):
    """z -> all pixels at once, ignoring coordinates (reference models.py:135-172).  Not the spatial
    decoder and not on the hot path: plain PyTorch modules, kept for CLI/API completeness."""

    def __init__(self, n, latent_dim, hidden_dim, n_out=1, num_layers=1, activation=nn.Tanh,
                 softplus=False, resid=False):
        super().__init__()
        self.n_out = n_out
        self.softplus = softplus
        mods = [nn.Linear(latent_dim, hidden_dim), activation()]
        for _ in range(num_layers - 1):
            mods += [nn.Linear(hidden_dim, hidden_dim), activation()]
        mods += [nn.Linear(hidden_dim, n * n_out), nn.Sigmoid()]
        if softplus:
            mods.append(nn.Softplus())
        self.layers = nn.Sequential(*mods)
        print(self)

    def forward(self, x, z):
        y = self.layers(z).view(z.size(0), -1, self.n_out)
        if self.softplus:
            y = torch.cat([F.softplus(y[:, :, :1]), y[:, :, 1:]], 2)
        return y
