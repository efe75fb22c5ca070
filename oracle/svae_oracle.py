"""CPU oracle for the spatial-VAE training-step hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package imports this file; only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may (as the checker / the timed CPU baseline, never as the product).

This is a functional restatement (plain tensors and closed formulas, no nn.Module) of
the arithmetic the reference performs with PyTorch modules.  Every function cites the
reference lines it follows (paths relative to the reference checkout):

  * decoder            spatial_vae/models.py:57-132   (SpatialGenerator)
  * encoder            spatial_vae/models.py:24-54    (InferenceNetwork)
  * step math, MNIST   train_mnist.py:24-90
  * step math, EM      train_particles.py:22-148
  * step math, galaxy  train_galaxy.py:27-128
  * pixel grid         train_mnist.py:316-320
  * CTF kernels        spatial_vae/ctf.py:7-56
  * Adam               torch.optim.Adam defaults as used at train_mnist.py:389-392

Parity pin: the reference ships no tests or golden vectors, so the
pins are outputs of the reference itself, generated in the build container by
``tests/golden/make_golden.py`` (imports the reference read-only) and committed as
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this file against them.

The arithmetic lives in a third-party dependency of the reference (PyTorch, pinned
``torch==1.7.1`` in requirements.txt:13; the container has 2.11).  The oracle uses only
elementwise torch ops, ``matmul`` and ``conv2d`` on CPU, in float32 or float64.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# configuration of one step (what the three reference drivers hard-code or take as flags)
# --------------------------------------------------------------------------------------

ACTIVATIONS = ("tanh", "leakyrelu", "relu", "sigmoid")


@dataclass
class StepConfig:
    """Knobs of eval_minibatch.  ``family`` picks the per-script quirks:

    mnist      theta-KL with mean penalty (train_mnist.py:62-63), Bernoulli, no z_scale
    particles  theta-KL without mean (train_particles.py:85-86), Gaussian (+fit-noise, CTF, mask)
    galaxy     theta-KL without mean (train_galaxy.py:97-99), Bernoulli over P*C
    """
    family: str = "mnist"
    rotate: bool = True
    translate: bool = True
    dx_scale: float = 0.1
    theta_prior: float = math.pi
    z_scale: float = 1.0
    activation: str = "tanh"
    softplus: bool = False
    resid: bool = False          # --resid: ResidLinear hidden layers in both networks (train_particles.py:291,437)


def act_fn(name: str):
    if name == "tanh":
        return torch.tanh
    if name == "leakyrelu":
        return lambda t: F.leaky_relu(t, 0.01)
    if name == "relu":
        return torch.relu
    if name == "sigmoid":
        return torch.sigmoid
    raise ValueError(name)


# --------------------------------------------------------------------------------------
# parameters as plain dictionaries of tensors
# --------------------------------------------------------------------------------------

def _linear_keys(sd, prefix="layers."):
    """(index, weight key, bias key) of every Linear inside a Sequential, plain or wrapped in ResidLinear
    (``layers.N.weight`` vs ``layers.N.linear.weight``, models.py:13-21,79-83)."""
    out = []
    for k in sd:
        if k.startswith(prefix) and k.endswith(".weight"):
            out.append((int(k.split(".")[1]), k, k[:-len("weight")] + "bias"))
    return sorted(out)


def decoder_params_from_state(sd: Dict[str, torch.Tensor]) -> Dict[str, object]:
    """state_dict of a SpatialGenerator -> plain dict.

    Key layout follows models.py:69-87: ``coord_linear``, ``latent_linear``, optional ``bilinear`` and the
    Sequential ``layers`` = [act, (Linear, act)*(L-1) | ResidLinear*(L-1), Linear, Sigmoid]; ``resid`` is True when
    the hidden layers are ResidLinear modules.
    """
    lin = _linear_keys(sd)
    hidden = [(sd[w], sd[b]) for _, w, b in lin[:-1]]
    out = {
        "coord_w": sd["coord_linear.weight"], "coord_b": sd["coord_linear.bias"],
        "latent_w": sd.get("latent_linear.weight"),
        "hidden": hidden,
        "out_w": sd[lin[-1][1]], "out_b": sd[lin[-1][2]],
    }
    if "bilinear.weight" in sd:
        out["bilinear_w"] = sd["bilinear.weight"]           # (H, coord features, Z)
    if any(".linear." in w for _, w, _ in lin):
        out["resid"] = True
    return out


def encoder_params_from_state(sd: Dict[str, torch.Tensor]) -> List[Tuple[torch.Tensor, torch.Tensor]]:
    """state_dict of an InferenceNetwork -> [(W, b)] in application order (models.py:31-43); ResidLinear layers
    (``layers.N.linear.*``) are returned like plain ones, pass resid=True to encoder_forward for them."""
    return [(sd[w], sd[b]) for _, w, b in _linear_keys(sd)]


def init_params(P_in: int, inf_dim: int, z_dim: int, H: int, L: int, Hq: int, Lq: int, C: int,
                seed: int = 0, dtype=torch.float32):
    """Random parameters with nn.Linear-like scale (uniform +-1/sqrt(fan_in)); own generator,
    NOT the reference's init stream (tests copy parameters explicitly)."""
    g = torch.Generator().manual_seed(seed)

    def lin(o, i, bias=True):
        k = 1.0 / math.sqrt(i)
        w = ((torch.rand(o, i, generator=g, dtype=torch.float64) * 2 - 1) * k).to(dtype)
        b = ((torch.rand(o, generator=g, dtype=torch.float64) * 2 - 1) * k).to(dtype) if bias else None
        return w, b

    cw, cb = lin(H, 2)
    lw = lin(H, z_dim, bias=False)[0] if z_dim > 0 else None
    hidden = [lin(H, H) for _ in range(L - 1)]
    ow, ob = lin(C, H)
    dec = {"coord_w": cw, "coord_b": cb, "latent_w": lw, "hidden": hidden, "out_w": ow, "out_b": ob}
    enc = [lin(Hq, P_in)] + [lin(Hq, Hq) for _ in range(Lq - 1)] + [lin(2 * inf_dim, Hq)]
    return dec, enc


def flatten_params(dec, enc) -> List[torch.Tensor]:
    """Parameter order used by the reference's optimiser: p_net.parameters() then
    q_net.parameters() (train_mnist.py:387), i.e. coord W,b; latent W; hidden W,b...; out W,b;
    then encoder W,b per layer."""
    out = [dec["coord_w"], dec["coord_b"]]
    if dec["latent_w"] is not None:
        out.append(dec["latent_w"])
    if dec.get("bilinear_w") is not None:
        out.append(dec["bilinear_w"])
    for w, b in dec["hidden"]:
        out += [w, b]
    out += [dec["out_w"], dec["out_b"]]
    for w, b in enc:
        out += [w, b]
    return out


def unflatten_like(dec, enc, flat: Sequence[torch.Tensor]):
    it = iter(flat)
    d = {"coord_w": next(it), "coord_b": next(it)}
    d["latent_w"] = next(it) if dec["latent_w"] is not None else None
    if dec.get("bilinear_w") is not None:
        d["bilinear_w"] = next(it)
    if dec.get("resid"):
        d["resid"] = True
    d["hidden"] = [(next(it), next(it)) for _ in dec["hidden"]]
    d["out_w"] = next(it)
    d["out_b"] = next(it)
    e = [(next(it), next(it)) for _ in enc]
    return d, e


# --------------------------------------------------------------------------------------
# pieces of the path
# --------------------------------------------------------------------------------------

def make_grid(n_rows: int, n_cols: int) -> torch.Tensor:
    """Pixel-centre coordinates (P,2), float32 (train_mnist.py:316-320): pixel p = i*m + j maps
    to (linspace(-1,1,m)[j], linspace(1,-1,n)[i]); built in float64 then cast."""
    xs = np.linspace(-1.0, 1.0, n_cols)
    ys = np.linspace(1.0, -1.0, n_rows)
    gx = np.tile(xs[None, :], (n_rows, 1)).reshape(-1)
    gy = np.repeat(ys, n_cols)
    return torch.from_numpy(np.stack([gx, gy], axis=1)).float()


def encoder_forward(enc, y: torch.Tensor, activation: str = "tanh", resid: bool = False):
    """InferenceNetwork.forward (models.py:46-54): Linear+act per hidden layer, final Linear,
    first half of the columns is z_mu, second half z_logstd.  resid: hidden layers after the first are
    ResidLinear, act(linear(x) + x) (models.py:13-21,35-36)."""
    a = act_fn(activation)
    h = y
    for i, (w, b) in enumerate(enc[:-1]):
        pre = h @ w.t() + b
        h = a(pre + h) if (resid and i > 0) else a(pre)
    w, b = enc[-1]
    o = h @ w.t() + b
    half = o.shape[1] // 2
    return o[:, :half], o[:, half:]


def transform_coords(grid: torch.Tensor, theta: Optional[torch.Tensor], dx: Optional[torch.Tensor]) -> torch.Tensor:
    """Rotate then translate the grid per image (train_mnist.py:50-59,70-74).
    bmm(x, [[c,s],[-s,c]]) gives x0' = x0*c - x1*s, x1' = x0*s + x1*c."""
    B = (theta if theta is not None else dx).shape[0]
    x = grid.unsqueeze(0).expand(B, -1, -1)
    if theta is not None:
        c, s = torch.cos(theta)[:, None], torch.sin(theta)[:, None]
        x = torch.stack([x[..., 0] * c - x[..., 1] * s, x[..., 0] * s + x[..., 1] * c], dim=-1)
    if dx is not None:
        x = x + dx[:, None, :]
    return x


def decoder_forward(dec, x: torch.Tensor, z: Optional[torch.Tensor], activation: str = "tanh",
                    softplus: bool = False) -> torch.Tensor:
    """SpatialGenerator.forward (models.py:90-132).  x (B,P,2), z (B,Z) -> (B,P,C) probabilities (the sigmoid is
    inside the module, :85).  Options follow the parameters present: 5 coordinate features when coord_w has 5
    columns (expand_coords, :99-102: x, x^2, x0*x1), a bilinear term when ``bilinear_w`` is given (:114-121),
    ResidLinear hidden layers when ``resid`` is set (:79-80)."""
    a = act_fn(activation)
    if dec["coord_w"].shape[1] == 5:                                 # :99-102
        x = torch.cat([x, x ** 2, (x[..., 0] * x[..., 1]).unsqueeze(-1)], dim=-1)
    h = x @ dec["coord_w"].t() + dec["coord_b"]                     # :104
    if dec["latent_w"] is not None and z is not None:
        h = h + (z @ dec["latent_w"].t())[:, None, :]                # :111-112,123
        if dec.get("bilinear_w") is not None:                        # nn.Bilinear(x, z): sum_ij x_i W[n,i,j] z_j
            h = h + torch.einsum("bpi,nij,bj->bpn", x, dec["bilinear_w"], z)
    h = a(h)                                                         # layers[0]
    for w, b in dec["hidden"]:
        pre = h @ w.t() + b
        h = a(pre + h) if dec.get("resid") else a(pre)
    y = torch.sigmoid(h @ dec["out_w"].t() + dec["out_b"])           # :84-85
    if softplus:                                                     # :129-130
        y = torch.cat([F.softplus(y[..., :1]), y[..., 1:]], dim=-1)
    return y


def bernoulli_loglik_per_image(y_hat: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """Per-image sum of the clamped Bernoulli log-likelihood.  F.binary_cross_entropy clamps
    each log term at -100 (train_mnist.py:80-81; mean * size == batch-mean of this sum)."""
    lp = torch.clamp(torch.log(y_hat), min=-100.0)
    lq = torch.clamp(torch.log(1.0 - y_hat), min=-100.0)
    return (y * lp + (1.0 - y) * lq).reshape(y.shape[0], -1).sum(1)


def ctf_correlate(mu: torch.Tensor, ctf: torch.Tensor, n: int) -> torch.Tensor:
    """Per-image cross-correlation of the decoded image with its own k x k kernel, zero padded
    to keep n x n (train_particles.py:112-119: grouped conv2d, no kernel flip)."""
    B = mu.shape[0]
    k = ctf.shape[-1]
    out = F.conv2d(mu.reshape(1, B, n, n), ctf.reshape(B, 1, ctf.shape[-2], k), padding=k // 2, groups=B)
    return out.reshape(B, n * n)


def step_forward(cfg: StepConfig, dec, enc, grid: torch.Tensor, y: torch.Tensor, eps: torch.Tensor,
                 ctf: Optional[torch.Tensor] = None, mask: Optional[torch.Tensor] = None,
                 theta_offset: Optional[torch.Tensor] = None, y_enc: Optional[torch.Tensor] = None):
    """One eval_minibatch.  y: (B,P) or (B,P,C) targets; eps: (B,I) the N(0,1) draw the
    reference takes from the global RNG (train_mnist.py:38); y_enc: what the encoder sees when
    augmentation rotated the input (train_particles.py:28-50), default y.

    Returns a dict with batch-mean ``elbo``, ``logp``, ``kl`` (the reference's return values)
    plus per-image ``logp_i``, ``kl_i`` and ``y_hat``.
    """
    B = y.shape[0]
    yin = (y if y_enc is None else y_enc).reshape(B, -1)
    z_mu, z_logstd = encoder_forward(enc, yin, cfg.activation, cfg.resid)      # train_mnist.py:32
    z_std = torch.exp(z_logstd)                                      # :33
    lat = z_std * eps + z_mu                                         # :39

    kl_i = torch.zeros(B, dtype=y.dtype)
    theta = None
    col = 0
    if cfg.rotate:                                                   # :42-63
        theta = lat[:, 0]
        t_mu, t_std, t_ls = z_mu[:, 0], z_std[:, 0], z_logstd[:, 0]
        s = cfg.theta_prior
        if cfg.family == "mnist":
            kl_i = -t_ls + math.log(s) + (t_std ** 2 + t_mu ** 2) / 2 / s ** 2 - 0.5
        else:                                                        # train_particles.py:85-86
            kl_i = -t_ls + math.log(s) + t_std ** 2 / 2 / s ** 2 - 0.5
        if theta_offset is not None:                                 # train_particles.py:71-74
            theta = theta + theta_offset
        col = 1
    dx = None
    zcol = col
    if cfg.translate:                                                # train_mnist.py:65-74
        dx = lat[:, col:col + 2] * cfg.dx_scale
        zcol = col + 2
    z = lat[:, zcol:]
    if cfg.family != "mnist":
        z = z * cfg.z_scale                                          # train_particles.py:99
    if theta is None and dx is None:
        x = grid.unsqueeze(0).expand(B, -1, -1)
    else:
        x = transform_coords(grid, theta, dx)
    y_hat = decoder_forward(dec, x, z if z.shape[1] > 0 else None, cfg.activation, cfg.softplus)

    P = grid.shape[0]
    if cfg.family in ("mnist", "galaxy"):
        logp_i = bernoulli_loglik_per_image(y_hat.reshape(B, -1), y.reshape(B, -1))
    else:
        params = y_hat.reshape(B, -1)                                # train_particles.py:102
        mu, logvar = params, None
        if params.shape[1] > P:                                      # fit-noise split :107-110
            mu, logvar = params[:, :P], params[:, P:]
        if ctf is not None:                                          # :112-119
            if logvar is not None:
                raise RuntimeError("CTF with fit-noise is rejected by the reference (train_particles.py:121-124,137)")
            mu = ctf_correlate(mu, ctf, int(round(math.sqrt(P))))
        yt = y.reshape(B, P)
        if mask is not None:                                         # :126-132
            yt, mu = yt[:, mask], mu[:, mask]
            if logvar is not None:
                logvar = logvar[:, mask]
        if logvar is not None:                                       # :136-139
            logp_i = -0.5 * ((mu - yt) ** 2 / torch.exp(logvar) + logvar).sum(1)
        else:
            logp_i = -0.5 * ((mu - yt) ** 2).sum(1)

    rest_mu, rest_std, rest_ls = z_mu[:, col:], z_std[:, col:], z_logstd[:, col:]
    kl_i = kl_i + (-rest_ls + 0.5 * rest_std ** 2 + 0.5 * rest_mu ** 2 - 0.5).sum(1)   # :84-85
    logp = logp_i.mean()
    kl = kl_i.mean()
    return {"elbo": logp - kl, "logp": logp, "kl": kl, "logp_i": logp_i, "kl_i": kl_i,
            "y_hat": y_hat, "z_mu": z_mu, "z_logstd": z_logstd, "latent": lat}


def step_grads(cfg: StepConfig, dec, enc, grid, y, eps, **kw):
    """Forward + autograd gradient of (-elbo) w.r.t. every parameter, in optimiser order."""
    flat = [p.detach().clone().requires_grad_(True) for p in flatten_params(dec, enc)]
    d, e = unflatten_like(dec, enc, flat)
    out = step_forward(cfg, d, e, grid, y, eps, **kw)
    (-out["elbo"]).backward()
    grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in flat]
    return {k: (v.detach() if torch.is_tensor(v) else v) for k, v in out.items()}, grads


@dataclass
class AdamState:
    """torch.optim.Adam(lr, betas=(0.9,0.999), eps=1e-8, weight_decay=0) restated."""
    lr: float = 1e-4
    b1: float = 0.9
    b2: float = 0.999
    eps: float = 1e-8
    t: int = 0
    m: List[torch.Tensor] = field(default_factory=list)
    v: List[torch.Tensor] = field(default_factory=list)

    def update(self, params: List[torch.Tensor], grads: List[torch.Tensor]) -> List[torch.Tensor]:
        if not self.m:
            self.m = [torch.zeros_like(p) for p in params]
            self.v = [torch.zeros_like(p) for p in params]
        self.t += 1
        out = []
        for i, (p, g) in enumerate(zip(params, grads)):
            self.m[i] = self.b1 * self.m[i] + (1 - self.b1) * g
            self.v[i] = self.b2 * self.v[i] + (1 - self.b2) * g * g
            bc1 = 1 - self.b1 ** self.t
            bc2 = 1 - self.b2 ** self.t
            denom = self.v[i].sqrt() / math.sqrt(bc2) + self.eps
            out.append(p - (self.lr / bc1) * self.m[i] / denom)
        return out


def train_steps(cfg: StepConfig, dec, enc, grid, batches, eps_list, lr=1e-4, **kw):
    """Run len(batches) reference train steps (eval_minibatch, backward, Adam.step) and return
    final (dec, enc) plus the per-step elbo list (train_mnist.py:138-150)."""
    adam = AdamState(lr=lr)
    elbos = []
    for y, eps in zip(batches, eps_list):
        out, grads = step_grads(cfg, dec, enc, grid, y, eps, **kw)
        new = adam.update(flatten_params(dec, enc), grads)
        dec, enc = unflatten_like(dec, enc, new)
        elbos.append(float(out["elbo"]))
    return dec, enc, elbos


# --------------------------------------------------------------------------------------
# CTF kernels (host precompute; the hot path consumes the result) -- spatial_vae/ctf.py
# --------------------------------------------------------------------------------------

def ctf_real_space_kernels(defocus_um, cs, voltage_kv, apix, bfactor, ampcont_pct, dfang_deg, n, m, scale=1.0):
    """Real-space CTF kernels (N,n,m) float32 = -fftshift(ifft2(CTF)).real (ctf.py:33-56).
    The reference passes ``defocus`` for both dfu and dfv (ctf.py:45-46), so the astigmatism
    term vanishes; wavelength / phase formulae per ctf.py:7-24."""
    fy = np.fft.fftfreq(n)
    fx = np.fft.fftfreq(m)
    a, b = np.meshgrid(fy, fx, indexing="ij")
    out = np.zeros((len(defocus_um), n, m), dtype=np.float32)
    for i in range(len(defocus_um)):
        ap = apix[i] * scale
        u, v = a.ravel() / ap, b.ravel() / ap
        volt = voltage_kv[i] * 1000.0
        csa = cs[i] * 1e7
        lam = 12.2639 / np.sqrt(volt + 0.97845e-6 * volt ** 2)
        s2 = u ** 2 + v ** 2
        df = defocus_um[i] * 10000.0
        gamma = 2 * np.pi * (-0.5 * df * lam * s2 + 0.25 * csa * lam ** 3 * s2 ** 2)
        w = ampcont_pct[i] / 100.0
        c = np.sqrt(1 - w ** 2) * np.sin(gamma) - w * np.cos(gamma)
        c = c * np.exp(-bfactor[i] / 4 * s2)
        out[i] = -np.fft.fftshift(np.fft.ifft2(c.reshape(n, m))).real
    return out
