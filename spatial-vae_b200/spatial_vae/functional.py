"""Host-side glue between PyTorch tensors and the C ABI (libsvae_b200.so).

Mirrors the step math of the reference drivers (train_mnist.py:24-90, train_particles.py:22-148,
train_galaxy.py:27-128) as ONE call, `elbo_step`, that the three `eval_minibatch` wrappers use, and
the module-level forwards (spatial_vae/models.py:46-54, 90-132) as autograd Functions.
PyTorch is plumbing here: it owns the device memory and the stream; all arithmetic runs in the
library.  There is no fallback: CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch
import torch.nn as nn

from . import _lib as L

_DEFAULT_PRECISION = os.environ.get("SVAE_PRECISION", "fast")


def default_precision() -> str:
    return _DEFAULT_PRECISION


def set_default_precision(p: str) -> None:
    global _DEFAULT_PRECISION
    if p not in L.PRECISION_CODES:
        raise ValueError(f"precision must be one of {sorted(L.PRECISION_CODES)}")
    _DEFAULT_PRECISION = p


def activation_code(act) -> int:
    """nn activation class / instance / name -> SVAE_ACT_* (LeakyReLU means slope 0.01)."""
    if isinstance(act, str):
        return L.ACT_CODES[act]
    cls = act if isinstance(act, type) else type(act)
    table = {nn.Tanh: L.ACT_TANH, nn.LeakyReLU: L.ACT_LEAKYRELU, nn.ReLU: L.ACT_RELU, nn.Sigmoid: L.ACT_SIGMOID}
    if cls not in table:
        raise NotImplementedError(f"activation {cls.__name__} is not implemented in the B200 kernels")
    return table[cls]


# ----------------------------------------------------------------------------------------------
# workspace + struct helpers
# ----------------------------------------------------------------------------------------------
_workspaces = {}


def workspace(nbytes: int, device: torch.device) -> torch.Tensor:
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        _workspaces.pop(key, None)
        ws = None
        ws = torch.empty(int(nbytes * 1.05) + 4096, dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("spatial_vae (B200 build) runs on CUDA tensors only: there is no CPU fallback. "
                               "Move the module and its inputs to a B200 device.")


def _f32(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


@dataclass
class DecoderTensors:
    coord_w: torch.Tensor
    coord_b: torch.Tensor
    latent_w: Optional[torch.Tensor]
    hidden: List[tuple]
    out_w: torch.Tensor
    out_b: torch.Tensor
    bilinear_w: Optional[torch.Tensor] = None      # (H, F, Z) nn.Bilinear weight (models.py:74-75)

    def flat(self) -> List[torch.Tensor]:
        """Parameter order of SpatialGenerator.parameters() (models.py:69-87)."""
        out = [self.coord_w, self.coord_b]
        if self.latent_w is not None:
            out.append(self.latent_w)
        if self.bilinear_w is not None:
            out.append(self.bilinear_w)
        for w, b in self.hidden:
            out += [w, b]
        out += [self.out_w, self.out_b]
        return out

    def layout(self) -> tuple:
        """(has_latent, n_hidden, has_bilinear): what from_flat needs to rebuild the structure."""
        return (self.latent_w is not None, len(self.hidden), self.bilinear_w is not None)

    @staticmethod
    def from_flat(flat: Sequence[torch.Tensor], has_latent: bool, n_hidden: int,
                  has_bilinear: bool = False) -> "DecoderTensors":
        it = iter(flat)
        cw, cb = next(it), next(it)
        lw = next(it) if has_latent else None
        bw = next(it) if has_bilinear else None
        hidden = [(next(it), next(it)) for _ in range(n_hidden)]
        return DecoderTensors(cw, cb, lw, hidden, next(it), next(it), bw)

    def struct(self) -> L.SvaeDecoderParams:
        s = L.SvaeDecoderParams()
        s.coord_w, s.coord_b = _ptr(self.coord_w), _ptr(self.coord_b)
        s.latent_w = _ptr(self.latent_w)
        for i, (w, b) in enumerate(self.hidden):
            s.hidden_w[i], s.hidden_b[i] = _ptr(w), _ptr(b)
        s.out_w, s.out_b = _ptr(self.out_w), _ptr(self.out_b)
        s.bilinear_w = _ptr(self.bilinear_w)
        return s


def encoder_struct(pairs: Sequence[tuple]) -> L.SvaeEncoderParams:
    s = L.SvaeEncoderParams()
    for i, (w, b) in enumerate(pairs):
        s.w[i], s.b[i] = _ptr(w), _ptr(b)
    return s


def _linears(seq) -> List[nn.Linear]:
    """The Linear modules of a reference-shaped Sequential, ResidLinear containers unwrapped (models.py:13-21)."""
    out = []
    for m in seq:
        if isinstance(m, nn.Linear):
            out.append(m)
        elif isinstance(getattr(m, "linear", None), nn.Linear):
            out.append(m.linear)
    return out


def decoder_tensors_of(p_net) -> DecoderTensors:
    """Pull the parameter tensors out of a SpatialGenerator-shaped module (models.py:69-87)."""
    lins = _linears(p_net.layers)
    hidden = [(m.weight, m.bias) for m in lins[:-1]]
    lw = p_net.latent_linear.weight if hasattr(p_net, "latent_linear") else None
    bw = p_net.bilinear.weight if hasattr(p_net, "bilinear") else None
    return DecoderTensors(p_net.coord_linear.weight, p_net.coord_linear.bias, lw, hidden, lins[-1].weight,
                          lins[-1].bias, bw)


def encoder_pairs_of(q_net) -> List[tuple]:
    return [(m.weight, m.bias) for m in _linears(q_net.layers)]


# ----------------------------------------------------------------------------------------------
# step specification
# ----------------------------------------------------------------------------------------------
@dataclass
class StepSpec:
    """Everything eval_minibatch needs besides tensors; `family` selects the per-script quirks
    (theta-KL mean penalty in mnist only, Gaussian likelihood in particles)."""
    family: str = "mnist"                 # mnist | particles | galaxy
    rotate: bool = True
    translate: bool = True
    dx_scale: float = 0.1
    theta_prior: float = math.pi
    z_scale: float = 1.0
    activation: int = L.ACT_TANH
    softplus: bool = False
    precision: str = "fast"
    chunk_images: int = 0
    resid: bool = False                   # --resid: ResidLinear hidden layers in both networks

    def config(self, C_out: int, grad_scale: float, dec: Optional["DecoderTensors"] = None) -> L.SvaeConfig:
        c = L.SvaeConfig()
        c.rotate, c.translate = int(bool(self.rotate)), int(bool(self.translate))
        if self.family == "particles":
            c.likelihood = L.LIK_GAUSS_FITNOISE if C_out == 2 else L.LIK_GAUSS
        else:
            c.likelihood = L.LIK_BERNOULLI
        c.theta_kl_mean = 1 if self.family == "mnist" else 0
        c.activation = self.activation
        c.precision = L.PRECISION_CODES[self.precision]
        c.softplus = int(bool(self.softplus))
        c.chunk_images = int(self.chunk_images)
        c.theta_prior, c.dx_scale = float(self.theta_prior), float(self.dx_scale)
        c.z_scale = float(self.z_scale) if self.family != "mnist" else 1.0
        c.grad_scale = float(grad_scale)
        c.resid = int(bool(self.resid))
        if dec is not None:        # the coordinate options are visible in the parameter shapes (models.py:65-75)
            c.expand_coords = int(dec.coord_w.shape[1] == 5)
            c.bilinear = int(dec.bilinear_w is not None)
        return c


def make_shape(B, P, C_out, Cin, Z, I, H, Lp, Hq, Lq, n_rows=0, n_cols=0, k_ctf=0) -> L.SvaeShape:
    s = L.SvaeShape()
    s.B, s.P, s.n_rows, s.n_cols = int(B), int(P), int(n_rows), int(n_cols)
    s.C, s.Cin, s.Z, s.I, s.H, s.L, s.Hq, s.Lq, s.k_ctf = int(C_out), int(Cin), int(Z), int(I), int(H), int(Lp), \
        int(Hq), int(Lq), int(k_ctf)
    return s


def shape_of(dec: DecoderTensors, enc: Sequence[tuple], B: int, P: int, Cin: int, spec: StepSpec,
             n_rows=0, n_cols=0, k_ctf=0) -> L.SvaeShape:
    H = dec.coord_w.shape[0]
    Z = 0 if dec.latent_w is None else dec.latent_w.shape[1]
    I = enc[-1][0].shape[0] // 2
    exp_I = Z + (1 if spec.rotate else 0) + (2 if spec.translate else 0)
    if I != exp_I:
        raise ValueError(f"inference network emits {I} latent dims but z_dim + rotate + translate = {exp_I}")
    return make_shape(B, P, dec.out_w.shape[0], Cin, Z, I, H, len(dec.hidden) + 1, enc[0][0].shape[0], len(enc) - 1,
                      n_rows, n_cols, k_ctf)


def run_step(spec: StepSpec, dec: DecoderTensors, enc: Sequence[tuple], grid: torch.Tensor, y: torch.Tensor,
             eps: Optional[torch.Tensor], *, y_enc=None, theta_offset=None, ctf=None, mask=None, grad_dec=None,
             grad_enc=None, grad_scale: Optional[float] = None, want_y_hat=False, want_latent=False, rng=None,
             decoder_grads_event: Optional["torch.cuda.Event"] = None, stats_sum: Optional[torch.Tensor] = None):
    """Enqueue one svae_step.  Returns (stats (B,3), y_hat or None, latent or None).
    eps None: the library draws it in the kernel from rng = (seed, step_counter (int32 device tensor), global index of
    this call's first image): Philox keyed on (seed, step, global image index), independent of how a minibatch is split
    across ranks.  decoder_grads_event: a torch.cuda.Event the library records where every decoder gradient of this
    call is final (before the encoder backward), for an early gradient exchange on another stream.  stats_sum: a
    contiguous fp32 device tensor of 4 elements that receives [sum logp, sum kl, sum elbo, 0] over the B images."""
    if eps is None and rng is None:
        raise ValueError("run_step needs eps or rng=(seed, step_tensor, image_offset)")
    _require_cuda(grid, y, dec.coord_w, enc[0][0], *([eps] if eps is not None else [rng[1]]))
    dev = y.device
    B, P = y.shape[0], grid.shape[0]
    Cin = y[0].numel() // P if B > 0 else max(1, enc[0][0].shape[1] // P)
    n = int(round(math.sqrt(P)))
    k_ctf = 0 if ctf is None else int(ctf.shape[-1])
    shape = shape_of(dec, enc, B, P, Cin, spec, n_rows=n if n * n == P else 0, n_cols=n if n * n == P else 0,
                     k_ctf=k_ctf)
    cfg = spec.config(shape.C, (1.0 / max(B, 1)) if grad_scale is None else grad_scale, dec)
    nbytes = C.c_size_t(0)
    L.check(L.lib.svae_workspace_bytes(C.byref(shape), C.byref(cfg), C.byref(nbytes)), "svae_workspace_bytes")
    if B == 0:   # a rank whose slice of a ragged last minibatch is empty: nothing to enqueue
        return (torch.zeros(0, 3, dtype=torch.float32, device=dev),
                torch.zeros(0, P, shape.C, dtype=torch.float32, device=dev) if want_y_hat else None,
                torch.zeros(0, shape.I, dtype=torch.float32, device=dev) if want_latent else None)
    ws = workspace(nbytes.value, dev)

    grid, y, eps = _f32(grid), _f32(y.reshape(B, P * Cin)), _f32(eps)
    if y.shape[1] != enc[0][0].shape[1]:
        raise ValueError(f"the inference network expects {enc[0][0].shape[1]} inputs per image, y has {y.shape[1]}")
    if eps is not None and tuple(eps.shape) != (B, shape.I):
        raise ValueError(f"eps must be ({B}, {shape.I}), got {tuple(eps.shape)}")
    if ctf is not None and tuple(ctf.shape[-2:]) != (k_ctf, k_ctf):
        raise ValueError("ctf kernels must be square")
    if mask is not None and mask.numel() != P:
        raise ValueError(f"mask must have {P} elements")
    y_enc = _f32(y_enc.reshape(B, P * Cin)) if y_enc is not None else None
    theta_offset = _f32(theta_offset)
    ctf = _f32(ctf)
    mask_u8 = mask.to(torch.uint8).contiguous() if mask is not None else None

    inp = L.SvaeStepInputs()
    inp.grid, inp.y, inp.y_enc, inp.theta_offset = _ptr(grid), _ptr(y), _ptr(y_enc), _ptr(theta_offset)
    inp.eps, inp.ctf, inp.mask = _ptr(eps), _ptr(ctf), _ptr(mask_u8)
    if eps is None:
        seed, step_t, image_offset = rng
        if step_t.dtype != torch.int32:
            raise ValueError("rng step counter must be an int32 device tensor")
        inp.rng_step, inp.rng_seed, inp.rng_image_offset = step_t.data_ptr(), int(seed) & (2 ** 64 - 1), int(image_offset)
    if decoder_grads_event is not None:
        if not decoder_grads_event.cuda_event:
            raise ValueError("decoder_grads_event has no CUDA handle yet: record it once before the first step")
        inp.decoder_grads_event = decoder_grads_event.cuda_event
    stats = torch.empty(B, 3, dtype=torch.float32, device=dev)
    y_hat = torch.empty(B, P, shape.C, dtype=torch.float32, device=dev) if want_y_hat else None
    latent = torch.empty(B, shape.I, dtype=torch.float32, device=dev) if want_latent else None
    out = L.SvaeStepOutputs()
    out.stats, out.y_hat, out.latent = _ptr(stats), _ptr(y_hat), _ptr(latent)
    if stats_sum is not None:
        if stats_sum.dtype != torch.float32 or stats_sum.numel() != 4 or not stats_sum.is_contiguous() or \
                stats_sum.device != dev:
            raise ValueError("stats_sum must be a contiguous fp32 tensor of 4 elements on the step's device")
        out.stats_sum = stats_sum.data_ptr()

    dstruct, estruct = dec.struct(), encoder_struct(enc)
    gd = grad_dec.struct() if grad_dec is not None else None
    ge = encoder_struct(grad_enc) if grad_enc is not None else None
    rc = L.lib.svae_step(C.byref(shape), C.byref(cfg), C.byref(dstruct), C.byref(estruct), C.byref(inp), C.byref(out),
                         C.byref(gd) if gd is not None else None, C.byref(ge) if ge is not None else None,
                         ws.data_ptr(), ws.numel(), _stream())
    L.check(rc, "svae_step")
    # keep the temporaries alive until the call is enqueued (same stream => safe afterwards)
    del grid, y, eps, y_enc, theta_offset, ctf, mask_u8
    return stats, y_hat, latent


class _FusedElbo(torch.autograd.Function):
    """elbo/logp/kl of one minibatch; the backward of -elbo is computed together with the forward
    (it is the same fused pass) and handed to autograd when the caller runs loss.backward()."""

    @staticmethod
    def forward(ctx, spec, n_dec, aux, *params):
        grid, y, eps, y_enc, theta_offset, ctf, mask, want_y_hat, need_grad, layout = aux
        dec_flat, enc_flat = params[:n_dec], params[n_dec:]
        dec = DecoderTensors.from_flat([p.detach() for p in dec_flat], *layout)
        enc = [(enc_flat[i].detach(), enc_flat[i + 1].detach()) for i in range(0, len(enc_flat), 2)]
        gd = ge = None
        if need_grad:
            gflat = [torch.zeros_like(p, dtype=torch.float32) for p in params]
            gd = DecoderTensors.from_flat(gflat[:n_dec], *layout)
            ge = [(gflat[n_dec + i], gflat[n_dec + i + 1]) for i in range(0, len(enc_flat), 2)]
        stats, y_hat, _ = run_step(spec, dec, enc, grid, y, eps, y_enc=y_enc, theta_offset=theta_offset, ctf=ctf,
                                   mask=mask, grad_dec=gd, grad_enc=ge, want_y_hat=want_y_hat)
        means = stats.mean(0) if stats.shape[0] > 0 else stats.new_zeros(3)
        ctx.grads = gflat if need_grad else None
        outs = (means[2].clone(), means[0].clone(), means[1].clone(),
                y_hat if y_hat is not None else stats.new_zeros(0), stats)
        ctx.mark_non_differentiable(*outs[1:])
        return outs

    @staticmethod
    def backward(ctx, g_elbo, *_):
        if ctx.grads is None:
            raise RuntimeError("elbo was computed without gradients (no_grad or no parameter requires grad)")
        # the library produced d(-elbo)/dparam
        grads = tuple(-g_elbo * g for g in ctx.grads)
        ctx.grads = None
        return (None, None, None) + grads


def elbo_step(spec: StepSpec, x_coord, y, p_net, q_net, *, eps=None, y_enc=None, theta_offset=None, ctf=None,
              mask=None, want_y_hat=False):
    """The body of eval_minibatch.  Returns (elbo, logp, kl, y_hat, per_image_stats); elbo carries
    autograd history w.r.t. the parameters of p_net and q_net."""
    if bool(getattr(p_net, "resid", False)) != bool(getattr(q_net, "resid", False)):
        raise NotImplementedError("the fused step takes one resid flag for both networks, as the reference's --resid "
                                  "does (train_particles.py:291,437-446)")
    if bool(getattr(p_net, "resid", False)) != bool(spec.resid):
        spec = StepSpec(**{**spec.__dict__, "resid": bool(getattr(p_net, "resid", False))})
    dec = decoder_tensors_of(p_net)
    enc = encoder_pairs_of(q_net)
    _require_cuda(x_coord, y, dec.coord_w, enc[0][0])
    B = y.shape[0]
    I = enc[-1][0].shape[0] // 2
    if eps is None:
        # same draw as the reference: (B, I) standard normals from the global generator of x's device
        eps = torch.empty(B, I, dtype=torch.float32, device=x_coord.device).normal_()
    dflat = dec.flat()
    eflat = [t for pair in enc for t in pair]
    # autograd Functions run their forward with grad mode off, so decide here whether the backward is wanted
    need_grad = torch.is_grad_enabled() and any(t.requires_grad for t in dflat + eflat)
    aux = (x_coord, y, eps, y_enc, theta_offset, ctf, mask, want_y_hat, need_grad, dec.layout())
    elbo, logp, kl, y_hat, stats = _FusedElbo.apply(spec, len(dflat), aux, *dflat, *eflat)
    return elbo, logp, kl, (y_hat if want_y_hat else None), stats


# ----------------------------------------------------------------------------------------------
# module-level forwards
# ----------------------------------------------------------------------------------------------
class _EncoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, act, x, *flat):
        _require_cuda(x, flat[0])
        pairs = [(flat[i].detach(), flat[i + 1].detach()) for i in range(0, len(flat), 2)]
        B, n_in = x.shape
        Hq, Lq, I2 = pairs[0][0].shape[0], len(pairs) - 1, pairs[-1][0].shape[0]
        shape = make_shape(B, n_in, 1, 1, 0, I2 // 2, 1, 1, Hq, Lq)
        xf = _f32(x.detach())
        out = torch.empty(B, I2, dtype=torch.float32, device=x.device)
        acts = torch.empty(max(Lq, 1), B, Hq, dtype=torch.float32, device=x.device)
        es = encoder_struct(pairs)
        L.check(L.lib.svae_encoder_forward(C.byref(shape), act, C.byref(es), xf.data_ptr(), out.data_ptr(),
                                           acts.data_ptr(), _stream()), "svae_encoder_forward")
        ctx.act, ctx.shape = act, shape
        ctx.save_for_backward(xf, acts, *[t for p in pairs for t in p])
        ctx.x_needs_grad = x.requires_grad
        return out

    @staticmethod
    def backward(ctx, g_out):
        xf, acts, *flat = ctx.saved_tensors
        pairs = [(flat[i], flat[i + 1]) for i in range(0, len(flat), 2)]
        shape = ctx.shape
        grads = [torch.zeros_like(t) for t in flat]
        gpairs = [(grads[i], grads[i + 1]) for i in range(0, len(grads), 2)]
        g = _f32(g_out).clone()
        wide = max(shape.Hq, 2 * shape.I)
        scratch = torch.empty(2, shape.B, wide, dtype=torch.float32, device=xf.device)
        gx = torch.empty_like(xf) if ctx.x_needs_grad else None
        es, gs = encoder_struct(pairs), encoder_struct(gpairs)
        L.check(L.lib.svae_encoder_backward(C.byref(shape), ctx.act, C.byref(es), xf.data_ptr(), acts.data_ptr(),
                                            g.data_ptr(), C.byref(gs), _ptr(gx), scratch.data_ptr(), _stream()),
                "svae_encoder_backward")
        return (None, gx) + tuple(grads)


def encoder_forward(q_net, x: torch.Tensor) -> torch.Tensor:
    pairs = encoder_pairs_of(q_net)
    flat = [t for p in pairs for t in p]
    act = q_net.activation_code
    if getattr(q_net, "resid", False):
        act |= L.ENC_RESID
    return _EncoderFn.apply(act, x, *flat)


class _DecoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, meta, x, z, *flat):
        act, softplus, precision, resid, layout = meta
        _require_cuda(x, flat[0])
        dec = DecoderTensors.from_flat([t.detach() for t in flat], *layout)
        B, P = x.shape[0], x.shape[1]
        H, Cn = dec.coord_w.shape[0], dec.out_w.shape[0]
        Z = 0 if dec.latent_w is None else dec.latent_w.shape[1]
        shape = make_shape(B, P, Cn, 1, Z, Z, H, len(dec.hidden) + 1, 1, 1)
        spec = StepSpec(rotate=False, translate=False, activation=act, softplus=softplus, precision=precision,
                        resid=resid)
        cfg = spec.config(Cn, 1.0, dec)
        nbytes = C.c_size_t(0)
        L.check(L.lib.svae_workspace_bytes(C.byref(shape), C.byref(cfg), C.byref(nbytes)), "svae_workspace_bytes")
        ws = workspace(nbytes.value, x.device)
        xf = _f32(x.detach())
        zf = _f32(z.detach()) if (z is not None and Z > 0) else None
        y = torch.empty(B, P, Cn, dtype=torch.float32, device=x.device)
        ds = dec.struct()
        L.check(L.lib.svae_decoder_forward(C.byref(shape), C.byref(cfg), C.byref(ds), xf.data_ptr(), _ptr(zf),
                                           y.data_ptr(), ws.data_ptr(), ws.numel(), _stream()), "svae_decoder_forward")
        ctx.meta, ctx.shape, ctx.cfg = meta, shape, cfg
        ctx.save_for_backward(xf, zf if zf is not None else xf.new_zeros(0), *[t.detach() for t in flat])
        ctx.needs = (x.requires_grad, z is not None and z.requires_grad)
        return y

    @staticmethod
    def backward(ctx, g_y):
        xf, zf, *flat = ctx.saved_tensors
        layout = ctx.meta[4]
        dec = DecoderTensors.from_flat(flat, *layout)
        shape, cfg = ctx.shape, ctx.cfg
        grads = [torch.zeros_like(t) for t in flat]
        gdec = DecoderTensors.from_flat(grads, *layout)
        nbytes = C.c_size_t(0)
        L.check(L.lib.svae_workspace_bytes(C.byref(shape), C.byref(cfg), C.byref(nbytes)), "svae_workspace_bytes")
        ws = workspace(nbytes.value, xf.device)
        gx = torch.empty_like(xf) if ctx.needs[0] else None
        gz = torch.zeros(shape.B, max(shape.Z, 1), dtype=torch.float32, device=xf.device) if shape.Z > 0 else None
        ds, gs = dec.struct(), gdec.struct()
        gyf = _f32(g_y)
        L.check(L.lib.svae_decoder_backward(C.byref(shape), C.byref(cfg), C.byref(ds), xf.data_ptr(),
                                            zf.data_ptr() if shape.Z > 0 else None, gyf.data_ptr(), C.byref(gs),
                                            _ptr(gx), _ptr(gz), ws.data_ptr(), ws.numel(), _stream()),
                "svae_decoder_backward")
        gz_out = gz[:, :shape.Z] if (gz is not None and ctx.needs[1]) else None
        return (None, gx, gz_out) + tuple(grads)


def decoder_forward(p_net, x: torch.Tensor, z: Optional[torch.Tensor]) -> torch.Tensor:
    dec = decoder_tensors_of(p_net)
    meta = (p_net.activation_code, bool(p_net.softplus), getattr(p_net, "precision", None) or default_precision(),
            bool(getattr(p_net, "resid", False)), dec.layout())
    return _DecoderFn.apply(meta, x, z, *dec.flat())


# ----------------------------------------------------------------------------------------------
# small utilities exposed to the drivers
# ----------------------------------------------------------------------------------------------
def adam_step(param: torch.Tensor, grad: torch.Tensor, m: torch.Tensor, v: torch.Tensor, lr: float, t: int,
              betas=(0.9, 0.999), eps=1e-8, zero_grad=True) -> None:
    """torch.optim.Adam.step() + zero_grad() over one flat fp32 buffer (train_mnist.py:149-150)."""
    _require_cuda(param, grad, m, v)
    L.check(L.lib.svae_adam_step(param.data_ptr(), grad.data_ptr(), m.data_ptr(), v.data_ptr(), param.numel(),
                                 lr, betas[0], betas[1], eps, t, int(zero_grad), _stream()), "svae_adam_step")


def adam_step_graph(param, grad, m, v, lr, t_dev, bias_corr_dev, betas=(0.9, 0.999), eps=1e-8, zero_grad=True) -> None:
    """adam_step driven by a device-resident step counter (t_dev int32[1] is incremented, the bias corrections
    go to bias_corr_dev float32[2]): no host-side state, so the pair can be captured in a CUDA graph."""
    _require_cuda(param, grad, m, v, t_dev, bias_corr_dev)
    L.check(L.lib.svae_adam_tick(t_dev.data_ptr(), bias_corr_dev.data_ptr(), betas[0], betas[1], _stream()),
            "svae_adam_tick")
    L.check(L.lib.svae_adam_step_graph(param.data_ptr(), grad.data_ptr(), m.data_ptr(), v.data_ptr(), param.numel(),
                                       lr, betas[0], betas[1], eps, bias_corr_dev.data_ptr(), int(zero_grad),
                                       _stream()), "svae_adam_step_graph")


def gather_rows(src: torch.Tensor, index: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """src[index] for a 2-D+ fp32 dataset resident on the GPU (replaces the per-sample DataLoader fetch)."""
    _require_cuda(src, index)
    if src.dtype != torch.float32:
        raise ValueError(f"gather_rows fetches fp32 datasets, got {src.dtype}")
    if not src.is_contiguous():
        raise ValueError("gather_rows needs a contiguous dataset tensor")
    src2 = src.reshape(src.shape[0], -1)
    idx = index.to(torch.int64).contiguous()
    if out is None:
        out = torch.empty((idx.numel(),) + tuple(src.shape[1:]), dtype=torch.float32, device=src.device)
    if idx.numel() == 0:          # empty tensors have a NULL data_ptr: nothing to enqueue
        return out
    L.check(L.lib.svae_gather_rows(src2.data_ptr(), idx.data_ptr(), out.data_ptr(), idx.numel(), src2.shape[1],
                                   _stream()), "svae_gather_rows")
    return out


_rot_staging = {}


def rotate_bicubic(y: torch.Tensor, n_rows: int, n_cols: int, angles_deg, channels: int = 1,
                   quantize_u8: bool = False) -> torch.Tensor:
    """Rotate every image of the minibatch y (B, n_rows*n_cols[, channels]) counter-clockwise by its angle
    (degrees), on the device, with the arithmetic of Pillow's Image.rotate(angle, resample=BICUBIC) -- what the
    reference does per image on the host for --augment-rotation (train_particles.py:39-43, train_galaxy.py:47-54).
    The destination->source affine matrices come from svae_rotation_matrices (host helper of the library), which
    builds them exactly as PIL/Image.py does."""
    _require_cuda(y)
    B = y.shape[0]
    if B == 0:                    # a rank whose slice of a ragged minibatch is empty
        return y.clone()
    ang = torch.as_tensor(angles_deg, dtype=torch.float64).contiguous()
    # a small ring of pinned staging buffers per batch size: the async copies of earlier calls may still be
    # queued behind GPU work while the host already fills the next set
    slot = _rot_staging.setdefault(B, {"i": 0, "bufs": [(torch.empty(B, 6, dtype=torch.float64).pin_memory(),
                                                         torch.empty(B, dtype=torch.int32).pin_memory(),
                                                         torch.cuda.Event()) for _ in range(8)]})
    slot["i"] = (slot["i"] + 1) % 8
    mats, modes, ev = slot["bufs"][slot["i"]]
    ev.synchronize()          # the copy that last used this set has finished (no-op on first use)
    L.check(L.lib.svae_rotation_matrices(ang.data_ptr(), B, n_rows, n_cols, mats.data_ptr(), modes.data_ptr()),
            "svae_rotation_matrices")
    src = _f32(y)
    out = torch.empty_like(src)
    mats_d, modes_d = mats.to(y.device, non_blocking=True), modes.to(y.device, non_blocking=True)
    ev.record()
    L.check(L.lib.svae_rotate_bicubic(src.data_ptr(), out.data_ptr(), mats_d.data_ptr(), modes_d.data_ptr(), B, n_rows,
                                      n_cols, channels, int(quantize_u8), _stream()), "svae_rotate_bicubic")
    return out.view_as(y)


def ctf_filter(ctf_params, n: int, m: int, scale: float = 1.0, device=None) -> torch.Tensor:
    """spatial_vae.ctf.ctf_filter on the device: (N, n, m) fp32 real-space kernels for a parsed CTF table (a pandas
    frame or anything indexable by the column names), all particles in one launch instead of the reference's Python
    loop over numpy ifft2 calls (reference ctf.py:33-56)."""
    import numpy as np
    from .ctf import CTF_COLUMNS
    table = np.stack([np.asarray(ctf_params[c], dtype=np.float64) for c in CTF_COLUMNS], axis=1)
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    params = torch.from_numpy(np.ascontiguousarray(table)).to(dev)
    _require_cuda(params)
    out = torch.empty(params.shape[0], n, m, dtype=torch.float32, device=dev)
    L.check(L.lib.svae_ctf_filter(_ptr(params), params.shape[0], int(n), int(m), float(scale), _ptr(out), _stream()),
            "svae_ctf_filter")
    return out


def gemm_bf16(mode: int, A: torch.Tensor, W: torch.Tensor, *, M: int, N: int, K: int, bias=None, aux=None,
              activation: int = L.ACT_TANH, out: torch.Tensor) -> torch.Tensor:
    """Raw access to the tcgen05 GEMM building block (tests)."""
    _require_cuda(A, W, out)
    L.check(L.lib.svae_gemm_bf16(mode, M, N, K, A.data_ptr(), A.stride(0), W.data_ptr(), W.stride(0), _ptr(bias),
                                 _ptr(aux), aux.stride(0) if aux is not None else 0, activation, out.data_ptr(),
                                 out.stride(0), _stream()), "svae_gemm_bf16")
    return out


def gemm_dx_moments(delta: torch.Tensor, W: torch.Tensor, *, H: int, grid: torch.Tensor, img: torch.Tensor,
                    coord_w: torch.Tensor, hz: torch.Tensor, P: int, activation: int = L.ACT_TANH) -> torch.Tensor:
    """Raw access to the fused tail of the decoder backward (tests): per-image column moments S (B, 3, Hp) of
    (delta W) .* act'(h_0) with h_0 recomputed from (grid, img, coord_w, hz); delta (rows, Hp) and W (Hp, Hp) bf16."""
    _require_cuda(delta, W, grid, img, coord_w, hz)
    rows, Hp = delta.shape
    S = torch.zeros(img.shape[0], 3, Hp, dtype=torch.float32, device=delta.device)
    L.check(L.lib.svae_gemm_dx_moments(rows, H, Hp, delta.data_ptr(), delta.stride(0), W.data_ptr(), W.stride(0),
                                       activation, grid.data_ptr(), img.data_ptr(), coord_w.data_ptr(), hz.data_ptr(),
                                       S.data_ptr(), P, _stream()), "svae_gemm_dx_moments")
    return S


def gemm_dw_top(h_top: torch.Tensor, h_prev: torch.Tensor, g_o: torch.Tensor, out_w: torch.Tensor, *, H: int,
                activation: int = L.ACT_TANH, want_delta: bool = True):
    """Raw access to the fused head of the decoder backward (tests): returns (dW, d_out_w, d_out_b, d_b, delta) for
    delta = (g_o out_w) .* act'(h_top), dW = delta^T h_prev, d_out_w = g_o^T h_top; h_* (rows, Hp) bf16."""
    _require_cuda(h_top, h_prev, g_o, out_w)
    rows, Hp = h_top.shape
    C = g_o.shape[1]
    dev = h_top.device
    dW = torch.zeros(H, H, device=dev)
    d_out_w = torch.zeros(C, H, device=dev)
    d_out_b = torch.zeros(C, device=dev)
    d_b = torch.zeros(H, device=dev)
    delta = torch.zeros(rows, Hp, dtype=torch.bfloat16, device=dev) if want_delta else None
    g_pad = torch.zeros(rows * C + 4, dtype=torch.float32, device=dev)     # the kernel copies g_o in 16-byte units
    g_pad[:rows * C] = g_o.reshape(-1)
    L.check(L.lib.svae_gemm_dw_top(rows, H, Hp, h_top.data_ptr(), h_prev.data_ptr(), activation, g_pad.data_ptr(), C,
                                   out_w.data_ptr(), d_out_w.data_ptr(), d_out_b.data_ptr(), d_b.data_ptr(),
                                   dW.data_ptr(), _ptr(delta), _stream()), "svae_gemm_dw_top")
    return dW, d_out_w, d_out_b, d_b, delta


class _ResidLinearFn(torch.autograd.Function):
    """act(x W^T + b + x) (reference models.py:20-21) on the library's fp32 GEMM with the skip connection in its epilogue."""

    @staticmethod
    def forward(ctx, act, x, w, b):
        _require_cuda(x, w, b)
        xf = _f32(x.reshape(-1, x.shape[-1]))
        wf, bf = _f32(w.detach()), _f32(b.detach())
        out = torch.empty_like(xf)
        L.check(L.lib.svae_resid_linear_forward(xf.data_ptr(), wf.data_ptr(), bf.data_ptr(), out.data_ptr(), xf.shape[0],
                                                xf.shape[1], act, _stream()), "svae_resid_linear_forward")
        ctx.act = act
        ctx.save_for_backward(xf, wf, out)
        return out.view(x.shape)

    @staticmethod
    def backward(ctx, g):
        xf, wf, out = ctx.saved_tensors
        gf = _f32(g.reshape(out.shape))
        g_pre, g_x = torch.empty_like(out), torch.empty_like(out)
        g_w, g_b = torch.zeros_like(wf), torch.zeros(wf.shape[0], dtype=torch.float32, device=wf.device)
        L.check(L.lib.svae_resid_linear_backward(xf.data_ptr(), wf.data_ptr(), out.data_ptr(), gf.data_ptr(),
                                                 g_pre.data_ptr(), g_x.data_ptr(), g_w.data_ptr(), g_b.data_ptr(),
                                                 xf.shape[0], xf.shape[1], ctx.act, _stream()), "svae_resid_linear_backward")
        return None, g_x.view(g.shape), g_w, g_b


def resid_linear(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, act: int) -> torch.Tensor:
    if w.shape[0] != w.shape[1] or x.shape[-1] != w.shape[1]:
        raise ValueError("ResidLinear needs a square weight matching the input width (x + linear(x))")
    return _ResidLinearFn.apply(act, x, w, b)
