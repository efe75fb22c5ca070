"""GPU parity tests: the CUDA path (through the C ABI, via spatial_vae.functional) against the
golden fixtures produced by the reference and against the CPU oracle on seeded inputs.

Tolerances (north star): per-image ELBO within 1e-3 relative; parameters within 1e-4 after 10
Adam steps.  PARITY precision (fp32 FFMA) is held to much tighter bounds; FAST precision (bf16
tcgen05 hidden GEMMs, fp32 accumulate) is held to the north-star ELBO tolerance.
"""
import contextlib
import io
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import svae_oracle as O
from tests.helpers import cfg_of, golden_grads, load_case, oracle_params


def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def _sf():
    import spatial_vae.functional as SF
    return SF


def _to_dev(dec, enc, dev):
    SF = _sf()
    d = SF.DecoderTensors(dec["coord_w"].to(dev), dec["coord_b"].to(dev),
                          dec["latent_w"].to(dev) if dec["latent_w"] is not None else None,
                          [(w.to(dev), b.to(dev)) for w, b in dec["hidden"]], dec["out_w"].to(dev), dec["out_b"].to(dev))
    e = [(w.to(dev), b.to(dev)) for w, b in enc]
    return d, e


def _zeros_like_params(d, e):
    SF = _sf()
    gd = SF.DecoderTensors(torch.zeros_like(d.coord_w), torch.zeros_like(d.coord_b),
                           torch.zeros_like(d.latent_w) if d.latent_w is not None else None,
                           [(torch.zeros_like(w), torch.zeros_like(b)) for w, b in d.hidden],
                           torch.zeros_like(d.out_w), torch.zeros_like(d.out_b))
    ge = [(torch.zeros_like(w), torch.zeros_like(b)) for w, b in e]
    return gd, ge


def _spec(cfg: O.StepConfig, precision, chunk=0):
    SF = _sf()
    from spatial_vae import _lib as L
    return SF.StepSpec(family=cfg.family, rotate=cfg.rotate, translate=cfg.translate, dx_scale=cfg.dx_scale,
                       theta_prior=cfg.theta_prior, z_scale=cfg.z_scale, activation=L.ACT_CODES[cfg.activation],
                       softplus=cfg.softplus, precision=precision, chunk_images=chunk)


def _run_cuda(cfg, dec, enc, grid, y, eps, precision, chunk=0, **kw):
    dev = _cuda()
    SF = _sf()
    d, e = _to_dev(dec, enc, dev)
    gd, ge = _zeros_like_params(d, e)
    kw = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in kw.items()}
    stats, y_hat, _ = SF.run_step(_spec(cfg, precision, chunk), d, e, grid.to(dev), y.to(dev), eps.to(dev),
                                  grad_dec=gd, grad_enc=ge, want_y_hat=True, **kw)
    torch.cuda.synchronize()
    grads = [g.cpu() for g in gd.flat()] + [t.cpu() for pair in ge for t in pair]
    return stats.cpu(), y_hat.cpu(), grads


def _golden_inputs(d):
    kw = {}
    t = lambda k: torch.from_numpy(d[k]).float()
    if "ctf" in d:
        kw["ctf"] = t("ctf")
    if "mask" in d:
        kw["mask"] = torch.from_numpy(d["mask"])
    if "theta_offset" in d:
        kw["theta_offset"] = t("theta_offset")
        kw["y_enc"] = t("y_enc")
    return t("grid"), t("y"), t("eps"), kw


CASES = [("mnist_rt", "mnist"), ("mnist_r", "mnist"), ("mnist_t", "mnist"), ("mnist_none", "mnist"),
         ("mnist_leaky_L3", "mnist"), ("particles_plain", "particles"), ("particles_fitnoise", "particles"),
         ("particles_ctf", "particles"), ("particles_mask", "particles"), ("particles_augment", "particles"),
         ("particles_zscale0", "particles"), ("galaxy_rgb", "galaxy")]


@pytest.mark.parametrize("name,family", CASES)
def test_step_matches_reference_golden_parity_precision(name, family):
    d = load_case(name)
    dec, enc = oracle_params(d)
    cfg = cfg_of(d, family)
    grid, y, eps, kw = _golden_inputs(d)
    stats, y_hat, grads = _run_cuda(cfg, dec, enc, grid, y, eps, "parity", **kw)
    np.testing.assert_allclose(float(stats[:, 2].mean()), float(d["elbo"]), rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(float(stats[:, 0].mean()), float(d["logp"]), rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(float(stats[:, 1].mean()), float(d["kl"]), rtol=2e-5, atol=2e-6)
    if "y_hat" in d:
        np.testing.assert_allclose(y_hat.reshape(d["y_hat"].shape).numpy(), d["y_hat"], rtol=1e-5, atol=1e-6)
    for i, (g, r) in enumerate(zip(grads, golden_grads(d))):
        np.testing.assert_allclose(g.numpy(), r.numpy(), rtol=5e-4, atol=5e-6, err_msg=f"{name} grad {i}")


@pytest.mark.parametrize("name,family", CASES)
def test_step_matches_reference_golden_fast_precision(name, family):
    d = load_case(name)
    dec, enc = oracle_params(d)
    cfg = cfg_of(d, family)
    grid, y, eps, kw = _golden_inputs(d)
    stats, y_hat, grads = _run_cuda(cfg, dec, enc, grid, y, eps, "fast", **kw)
    # north-star tolerance on the ELBO: 1e-3 relative
    assert abs(float(stats[:, 2].mean()) - float(d["elbo"])) <= 1e-3 * abs(float(d["elbo"])) + 1e-4
    for i, (g, r) in enumerate(zip(grads, golden_grads(d))):
        scale = float(r.abs().max()) + 1e-6
        # (leaky)relu has a discontinuous derivative: bf16 rounding flips it for pre-activations near 0
        gtol = 3e-2 if cfg.activation == "tanh" else 1.5e-1
        assert float((g - r).abs().max()) <= gtol * scale, f"{name} grad {i}"


def _random_case(family, B, n, H, L, Z, Hq, C=1, seed=0, dtype=torch.float32):
    P = n * n
    Cin = 3 if family == "galaxy" else 1
    dec, enc = O.init_params(P * Cin, Z + 3, Z, H, L, Hq, 2, C, seed=seed)
    g = torch.Generator().manual_seed(1234 + seed)
    if family == "mnist":
        y = (torch.rand(B, P, generator=g) > 0.8).float() * torch.rand(B, P, generator=g)
    elif family == "galaxy":
        y = torch.rand(B, P, 3, generator=g)
    else:
        y = torch.randn(B, P, generator=g)
    eps = torch.randn(B, Z + 3, generator=g)
    return dec, enc, O.make_grid(n, n), y, eps


@pytest.mark.parametrize("precision,tol", [("parity", 2e-5), ("fast", 1e-3)])
def test_c1_shape_per_image_elbo_against_oracle(precision, tol):
    """BASELINE configs[0] shape (28x28, z=2, 500x2) at B=24: per-image ELBO vs the oracle."""
    dec, enc, grid, y, eps = _random_case("mnist", 24, 28, 500, 2, 2, 500)
    cfg = O.StepConfig(family="mnist", theta_prior=math.pi / 4)
    out, ograds = O.step_grads(cfg, dec, enc, grid, y, eps)
    stats, y_hat, grads = _run_cuda(cfg, dec, enc, grid, y, eps, precision)
    ref = (out["logp_i"] - out["kl_i"]).numpy()
    rel = np.abs(stats[:, 2].numpy() - ref) / np.abs(ref)
    assert rel.max() <= tol, f"max per-image ELBO rel err {rel.max():.3e}"
    gtol = 1e-3 if precision == "parity" else 5e-2
    for i, (g, r) in enumerate(zip(grads, ograds)):
        scale = float(r.abs().max()) + 1e-8
        assert float((g - r).abs().max()) <= gtol * scale, f"grad {i}: {float((g - r).abs().max()) / scale:.3e}"


@pytest.mark.parametrize("precision", ["parity", "fast"])
def test_galaxy_shape_rgb_L4(precision):
    """galaxy-like: RGB targets, 4-layer decoder, H not a multiple of 64."""
    dec, enc, grid, y, eps = _random_case("galaxy", 6, 16, 200, 4, 5, 96, C=3, seed=3)
    cfg = O.StepConfig(family="galaxy", theta_prior=math.pi, z_scale=1.0)
    out, ograds = O.step_grads(cfg, dec, enc, grid, y, eps)
    stats, y_hat, grads = _run_cuda(cfg, dec, enc, grid, y, eps, precision)
    ref = (out["logp_i"] - out["kl_i"]).numpy()
    tol = 2e-5 if precision == "parity" else 1e-3
    assert (np.abs(stats[:, 2].numpy() - ref) / np.abs(ref)).max() <= tol
    np.testing.assert_allclose(y_hat.numpy(), out["y_hat"].numpy(), atol=1e-5 if precision == "parity" else 5e-3)


@pytest.mark.parametrize("precision", ["parity", "fast"])
def test_particles_ctf_40x40(precision):
    """C5-like: 40x40 particles with 39x39 CTF kernels."""
    B, n = 5, 40
    dec, enc, grid, y, eps = _random_case("particles", B, n, 128, 2, 2, 64, seed=5)
    g = torch.Generator().manual_seed(9)
    ctf = 0.03 * torch.randn(B, 1, 39, 39, generator=g)
    cfg = O.StepConfig(family="particles", theta_prior=math.pi)
    out, ograds = O.step_grads(cfg, dec, enc, grid, y, eps, ctf=ctf)
    stats, _, grads = _run_cuda(cfg, dec, enc, grid, y, eps, precision, ctf=ctf)
    ref = (out["logp_i"] - out["kl_i"]).numpy()
    tol = 2e-5 if precision == "parity" else 1e-3
    assert (np.abs(stats[:, 2].numpy() - ref) / np.abs(ref)).max() <= tol
    gtol = 1e-3 if precision == "parity" else 5e-2
    for i, (gg, r) in enumerate(zip(grads, ograds)):
        scale = float(r.abs().max()) + 1e-8
        assert float((gg - r).abs().max()) <= gtol * scale, f"grad {i}"


def test_chunking_and_batch_split_are_invariant():
    """Processing the minibatch in chunks, or as two half-batches with grad_scale = 1/B (the
    data-parallel decomposition), gives the same per-image stats and the same summed gradient."""
    dec, enc, grid, y, eps = _random_case("mnist", 10, 12, 96, 3, 4, 48, seed=7)
    cfg = O.StepConfig(family="mnist", theta_prior=math.pi / 4)
    s0, _, g0 = _run_cuda(cfg, dec, enc, grid, y, eps, "parity")
    s1, _, g1 = _run_cuda(cfg, dec, enc, grid, y, eps, "parity", chunk=3)
    np.testing.assert_allclose(s0.numpy(), s1.numpy(), rtol=1e-6, atol=1e-6)
    for a, b in zip(g0, g1):   # summation order differs between the two schedules: fp32 noise only
        np.testing.assert_allclose(a.numpy(), b.numpy(), rtol=1e-4, atol=1e-5 * float(a.abs().max()))
    sa, _, ga = _run_cuda(cfg, dec, enc, grid, y[:6], eps[:6], "parity", grad_scale=0.1)
    sb, _, gb = _run_cuda(cfg, dec, enc, grid, y[6:], eps[6:], "parity", grad_scale=0.1)
    np.testing.assert_allclose(torch.cat([sa, sb]).numpy(), s0.numpy(), rtol=1e-6, atol=1e-6)
    for a, b, c in zip(ga, gb, g0):
        np.testing.assert_allclose((a + b).numpy(), c.numpy(), rtol=1e-4, atol=1e-5 * float(c.abs().max()))


def test_empty_batch_is_a_noop():
    dec, enc, grid, y, eps = _random_case("mnist", 2, 8, 32, 2, 2, 16, seed=8)
    cfg = O.StepConfig(family="mnist")
    stats, y_hat, grads = _run_cuda(cfg, dec, enc, grid, y[:0], eps[:0], "parity")
    assert stats.shape == (0, 3) and all(float(g.abs().max()) == 0.0 for g in grads)


# ---- the tcgen05 GEMM building block -------------------------------------------------------------
@pytest.mark.parametrize("M,N,K", [(1000, 512, 512), (128, 64, 64), (4096 + 77, 1024, 1024), (300, 192, 320)])
def test_tc_gemm_forward(M, N, K):
    dev = _cuda()
    SF = _sf()
    g = torch.Generator().manual_seed(M + N)
    A = (torch.randn(M, K, generator=g) * 0.5).to(dev).bfloat16()
    W = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(dev).bfloat16()
    bias = torch.randn(N, generator=g).to(dev)
    out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
    SF.gemm_bf16(0, A, W, M=M, N=N, K=K, bias=bias, activation=0, out=out)
    ref = torch.tanh(A.float() @ W.float().t() + bias)
    torch.cuda.synchronize()
    assert float((out.float() - ref).abs().max()) < 1.5e-2


@pytest.mark.parametrize("M,N,K", [(1000, 512, 512), (128, 64, 64), (2048 + 5, 1024, 1024), (300, 192, 320)])
def test_tc_gemm_dx(M, N, K):
    dev = _cuda()
    SF = _sf()
    g = torch.Generator().manual_seed(M + N + 1)
    A = (torch.randn(M, K, generator=g) * 0.5).to(dev).bfloat16()
    W = (torch.randn(K, N, generator=g) / math.sqrt(K)).to(dev).bfloat16()
    aux = torch.tanh(torch.randn(M, N, generator=g)).to(dev).bfloat16()
    out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
    SF.gemm_bf16(1, A, W, M=M, N=N, K=K, aux=aux, activation=0, out=out)
    ref = (A.float() @ W.float()) * (1 - aux.float() ** 2)
    torch.cuda.synchronize()
    assert float((out.float() - ref).abs().max()) < 1.5e-2


@pytest.mark.parametrize("M,N,K,ld", [(500, 500, 5000, 512), (512, 512, 78400, 512), (1000, 1000, 4099, 1024),
                                      (30, 30, 700, 64)])
def test_tc_gemm_dw(M, N, K, ld):
    dev = _cuda()
    SF = _sf()
    g = torch.Generator().manual_seed(M + K)
    A = torch.zeros(K, ld, dtype=torch.bfloat16, device=dev)
    Bm = torch.zeros(K, ld, dtype=torch.bfloat16, device=dev)
    A[:, :M] = (torch.randn(K, M, generator=g) * 0.1).to(dev).bfloat16()
    Bm[:, :N] = (torch.randn(K, N, generator=g) * 0.1).to(dev).bfloat16()
    out = torch.ones(M, N, device=dev, dtype=torch.float32)
    SF.gemm_bf16(2, A, Bm, M=M, N=N, K=K, out=out)
    ref = 1.0 + A[:, :M].float().t() @ Bm[:, :N].float()
    torch.cuda.synchronize()
    err = float((out - ref).abs().max())
    assert err < 2e-3 * max(1.0, float(ref.abs().max())), err


# rows = B * P; (B, P, H, Hp): several images per 256-row tile, images spanning tiles, ragged last tile, padded width
@pytest.mark.parametrize("B,P,H,Hp,act", [(3, 300, 500, 512, 0), (40, 16, 60, 64, 0), (2, 784, 500, 512, 1),
                                          (5, 100, 1000, 1024, 0), (7, 37, 130, 192, 3)])
def test_tc_dx_moments(B, P, H, Hp, act):
    """Transposed dX GEMM with h_0 recomputed and delta_0 reduced per image in the epilogue (tc_bwd.cu) against
    plain fp32 torch: S[b, {1, c0, c1}, n] = sum_p ((delta W) .* act'(h_0))[b, p, n] {1, grid[p]}."""
    dev = _cuda()
    SF = _sf()
    from spatial_vae import _lib as L
    if L.lib.svae_device_sm_count() < 2 * ((Hp + 255) // 256):
        pytest.skip("one CTA pair per 256-column tile is the kernel's minimum (small emulated devices)")
    g = torch.Generator().manual_seed(B * 1000 + P)
    rows = B * P
    delta = torch.zeros(rows, Hp, dtype=torch.bfloat16)
    delta[:, :H] = (torch.randn(rows, H, generator=g) * 0.1).bfloat16()
    W = torch.zeros(Hp, Hp, dtype=torch.bfloat16)
    W[:H, :H] = (torch.randn(H, H, generator=g) / math.sqrt(H)).bfloat16()
    grid = torch.rand(P, 2, generator=g) * 2 - 1
    theta = torch.randn(B, generator=g)
    img = torch.stack([torch.cos(theta), torch.sin(theta), 0.1 * torch.randn(B, generator=g),
                       0.1 * torch.randn(B, generator=g)], 1).contiguous()
    coord_w = torch.randn(H, 2, generator=g)
    hz = torch.full((B, Hp), float("nan"))
    hz[:, :H] = torch.randn(B, H, generator=g)
    S = SF.gemm_dx_moments(delta.to(dev), W.to(dev), H=H, grid=grid.to(dev), img=img.to(dev), coord_w=coord_w.to(dev),
                           hz=hz.to(dev), P=P, activation=act)
    xp = grid[None, :, 0] * img[:, None, 0] - grid[None, :, 1] * img[:, None, 1] + img[:, None, 2]
    yp = grid[None, :, 0] * img[:, None, 1] + grid[None, :, 1] * img[:, None, 0] + img[:, None, 3]
    a0 = xp[..., None] * coord_w[:, 0] + yp[..., None] * coord_w[:, 1] + hz[:, None, :H]
    if act == 0:
        h0 = torch.tanh(a0); dact = 1 - h0 * h0
    elif act == 1:
        dact = torch.where(a0 > 0, 1.0, 0.01)
    else:
        h0 = torch.sigmoid(a0); dact = h0 * (1 - h0)
    d0 = (delta.float() @ W.float())[:, :H].view(B, P, H) * dact
    ref = torch.stack([d0.sum(1), (d0 * grid[None, :, 0, None]).sum(1), (d0 * grid[None, :, 1, None]).sum(1)], 1)
    torch.cuda.synchronize()
    got = S.cpu()[:, :, :H]
    err = float((got - ref).abs().max())
    assert err < 3e-3 * max(1.0, float(ref.abs().max())), err
    assert torch.isfinite(S.cpu()[:, :, :H]).all()


@pytest.mark.parametrize("rows,H,Hp,C,act", [(3000, 500, 512, 1, 0), (777, 60, 64, 2, 0), (5000, 1000, 1024, 3, 0),
                                             (1500, 130, 192, 1, 1)])
def test_tc_dw_top(rows, H, Hp, C, act):
    """Top-layer weight-gradient GEMM with delta built in shared memory from h_top and g_o (tc_bwd.cu) against torch."""
    dev = _cuda()
    SF = _sf()
    g = torch.Generator().manual_seed(rows + H)
    h_top = torch.zeros(rows, Hp, dtype=torch.bfloat16)
    h_prev = torch.zeros(rows, Hp, dtype=torch.bfloat16)
    h_top[:, :H] = torch.tanh(torch.randn(rows, H, generator=g)).bfloat16()
    h_prev[:, :H] = torch.tanh(torch.randn(rows, H, generator=g)).bfloat16()
    g_o = torch.randn(rows, C, generator=g) * 0.1
    out_w = torch.randn(C, H, generator=g) / math.sqrt(H)
    dW, d_out_w, d_out_b, d_b, delta = SF.gemm_dw_top(h_top.to(dev), h_prev.to(dev), g_o.to(dev), out_w.to(dev), H=H,
                                                      activation=act)
    ht = h_top.float()[:, :H]
    dact = 1 - ht * ht if act == 0 else torch.where(ht > 0, 1.0, 0.01)
    d_ref = (g_o @ out_w) * dact
    torch.cuda.synchronize()
    dq = delta.cpu().float()[:, :H]
    assert float((dq - d_ref).abs().max()) < 1e-2 * float(d_ref.abs().max())
    assert float(delta.cpu().float()[:, H:].abs().max() if Hp > H else 0.0) == 0.0
    ref_dW = dq.t() @ h_prev.float()[:, :H]            # the GEMM consumes the bf16-rounded delta
    scale = max(1.0, float(ref_dW.abs().max()))
    assert float((dW.cpu() - ref_dW).abs().max()) < 2e-3 * scale
    assert float((d_b.cpu() - d_ref.sum(0)).abs().max()) < 1e-3 * max(1.0, float(d_ref.sum(0).abs().max()))
    ref_ow = g_o.t() @ ht
    assert float((d_out_w.cpu() - ref_ow).abs().max()) < 1e-3 * max(1.0, float(ref_ow.abs().max()))
    assert float((d_out_b.cpu() - g_o.sum(0)).abs().max()) < 1e-3 * max(1.0, float(g_o.sum(0).abs().max()))


# ---- module-level API (spatial_vae.models) --------------------------------------------------------
def _modules_from_golden(d, dev, C=1):
    import spatial_vae.models as M
    import torch.nn as nn
    p_state = {k[2:]: torch.from_numpy(v) for k, v in d.items() if k.startswith("p.")}
    H = p_state["coord_linear.weight"].shape[0]
    Z = p_state["latent_linear.weight"].shape[1]
    Lp = int(d["L"])
    n_out = [v for k, v in p_state.items() if k.endswith(".weight")][-1].shape[0]
    with contextlib.redirect_stdout(io.StringIO()):
        p = M.SpatialGenerator(Z, H, n_out=n_out, num_layers=Lp, activation=nn.Tanh)
    p.load_state_dict(p_state)
    return p.to(dev)


@pytest.mark.parametrize("precision", ["parity", "fast"])
def test_spatial_generator_module_forward_backward(precision):
    dev = _cuda()
    d = load_case("decoder_module")
    p = _modules_from_golden(d, dev)
    p.precision = precision
    x = torch.from_numpy(d["x"]).to(dev).requires_grad_(True)
    z = torch.from_numpy(d["z"]).to(dev).requires_grad_(True)
    y = p(x, z)
    tol = 1e-5 if precision == "parity" else 5e-3
    np.testing.assert_allclose(y.detach().cpu().numpy(), d["y_hat"], atol=tol)
    w = torch.linspace(-1, 1, y.numel(), device=dev).reshape(y.shape)
    (y * w).sum().backward()
    # oracle gradient by autograd on the CPU restatement
    from tests.helpers import state
    dec = O.decoder_params_from_state(state(d, "p."))
    flat = [t.clone().requires_grad_(True) for t in O.flatten_params(dec, [])]
    dd, _ = O.unflatten_like(dec, [], flat)
    xc = torch.from_numpy(d["x"]).requires_grad_(True)
    zc = torch.from_numpy(d["z"]).requires_grad_(True)
    yo = O.decoder_forward(dd, xc, zc)
    (yo * w.cpu()).sum().backward()
    gtol = 1e-3 if precision == "parity" else 5e-2

    def close(a, b, what):
        scale = float(b.abs().max()) + 1e-8
        assert float((a.cpu() - b).abs().max()) <= gtol * scale, what

    close(x.grad, xc.grad, "g_x")
    close(z.grad, zc.grad, "g_z")
    for (name, prm), ref in zip(p.named_parameters(), flat):
        close(prm.grad, ref.grad, name)


def test_inference_network_module_forward_backward():
    dev = _cuda()
    import spatial_vae.models as M
    import torch.nn as nn
    d = load_case("mnist_rt")
    q_state = {k[2:]: torch.from_numpy(v) for k, v in d.items() if k.startswith("q.")}
    n_in = q_state["layers.0.weight"].shape[1]
    Hq = q_state["layers.0.weight"].shape[0]
    I = q_state["layers.4.weight"].shape[0] // 2
    with contextlib.redirect_stdout(io.StringIO()):
        q = M.InferenceNetwork(n_in, I, Hq, num_layers=2, activation=nn.Tanh)
    q.load_state_dict(q_state)
    q = q.to(dev)
    y = torch.from_numpy(d["y"]).to(dev).requires_grad_(True)
    mu, ls = q(y)
    enc = O.encoder_params_from_state(q_state)
    flat = [t.clone().requires_grad_(True) for pair in enc for t in pair]
    yc = torch.from_numpy(d["y"]).requires_grad_(True)
    mo, lo = O.encoder_forward([(flat[i], flat[i + 1]) for i in range(0, len(flat), 2)], yc)
    np.testing.assert_allclose(mu.detach().cpu().numpy(), mo.detach().numpy(), rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(ls.detach().cpu().numpy(), lo.detach().numpy(), rtol=1e-4, atol=1e-6)
    ((mu ** 2).sum() + ls.sum()).backward()
    ((mo ** 2).sum() + lo.sum()).backward()
    np.testing.assert_allclose(y.grad.cpu().numpy(), yc.grad.numpy(), rtol=1e-3, atol=1e-6)
    for prm, ref in zip(q.parameters(), flat):
        np.testing.assert_allclose(prm.grad.cpu().numpy(), ref.grad.numpy(), rtol=1e-3, atol=1e-6)


# ---- 10-step Adam trajectory -------------------------------------------------------------------------
@pytest.mark.parametrize("precision", ["parity", "parity_tc", "fast"])
def test_adam_trajectory_matches_reference(precision):
    """Reference loop (eval_minibatch, backward, Adam.step, zero_grad) x10 on the golden fixture (written by the
    reference itself): parameters within 1e-4 (north star) for PARITY (fp32 FFMA) AND for PARITY_TC (3-term bf16
    splits on tcgen05).  FAST (single bf16 pass) cannot meet the max-abs gate -- Adam turns relative gradient noise
    on near-zero gradients into O(lr) differences -- so it is held to what bounds a trajectory: the
    fraction of parameters within 1e-4, the mean absolute difference, and the relative L2 error of the total
    10-step update (test_adam_trajectory_at_c1_shape); on THIS fixture all three modes are inside 1e-4.  The
    measured values are written to gpurun_out/parity_trajectory_<precision>.json."""
    dev = _cuda()
    SF = _sf()
    d = load_case("mnist_adam10")
    dec, enc = oracle_params(d, p_prefix="init.p.", q_prefix="init.q.")
    dd, ee = _to_dev(dec, enc, dev)
    params = dd.flat() + [t for pair in ee for t in pair]
    flat = torch.cat([p.reshape(-1) for p in params]).clone()
    # re-point the parameter tensors at views of one flat buffer (what the trainer does)
    views, off = [], 0
    for p in params:
        views.append(flat[off:off + p.numel()].view_as(p))
        off += p.numel()
    n_dec = len(dd.flat())
    dd = SF.DecoderTensors.from_flat(views[:n_dec], dd.latent_w is not None, len(dd.hidden))
    ee = [(views[n_dec + i], views[n_dec + i + 1]) for i in range(0, len(views) - n_dec, 2)]
    gflat = torch.zeros_like(flat)
    gviews, off = [], 0
    for p in params:
        gviews.append(gflat[off:off + p.numel()].view_as(p))
        off += p.numel()
    gd = SF.DecoderTensors.from_flat(gviews[:n_dec], dd.latent_w is not None, len(dd.hidden))
    ge = [(gviews[n_dec + i], gviews[n_dec + i + 1]) for i in range(0, len(gviews) - n_dec, 2)]
    m, v = torch.zeros_like(flat), torch.zeros_like(flat)
    cfg = O.StepConfig(family="mnist", theta_prior=float(d["theta_prior"]), dx_scale=float(d["dx_scale"]))
    grid = torch.from_numpy(d["grid"]).to(dev)
    elbos = []
    for t in range(d["ys"].shape[0]):
        y = torch.from_numpy(d["ys"][t]).to(dev)
        eps = torch.from_numpy(d["eps"][t]).to(dev)
        stats, _, _ = SF.run_step(_spec(cfg, precision), dd, ee, grid, y, eps, grad_dec=gd, grad_enc=ge)
        SF.adam_step(flat, gflat, m, v, float(d["lr"]), t + 1)
        elbos.append(float(stats[:, 2].mean()))
    fdec, fenc = oracle_params(d, p_prefix="final.p.", q_prefix="final.q.")
    ref = torch.cat([p.reshape(-1) for p in O.flatten_params(fdec, fenc)])
    diff = (flat.cpu() - ref).abs()
    np.testing.assert_allclose(elbos, d["elbos"], rtol=1e-4 if precision != "fast" else 1e-3)
    idec, ienc = oracle_params(d, p_prefix="init.p.", q_prefix="init.q.")
    init = torch.cat([p.reshape(-1) for p in O.flatten_params(idec, ienc)])
    frac = float((diff < 1e-4).float().mean())
    rel_update = float((flat.cpu() - ref).norm() / (ref - init).norm())
    rec = {"precision": precision, "max_abs_dparam": float(diff.max()), "mean_abs_dparam": float(diff.mean()),
           "frac_within_1e-4": frac, "rel_l2_error_of_10_step_update": rel_update,
           "max_rel_elbo_err": float(np.max(np.abs(np.array(elbos) - d["elbos"]) / np.abs(d["elbos"])))}
    print(rec)
    try:
        import json
        os.makedirs(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out"), exist_ok=True)
        with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out",
                               f"parity_trajectory_{precision}.json"), "w") as f:
            json.dump(rec, f)
    except OSError:
        pass
    # every precision meets the north-star gate on this fixture (measured on a B200: max |dparam| 1.5e-8 PARITY,
    # 3.0e-8 PARITY_TC, 6.2e-6 FAST); the C1-shape trajectory below is the harder case for FAST
    assert float(diff.max()) < 1e-4, rec
    if precision == "fast":
        assert rel_update < 0.01, rec


@pytest.mark.parametrize("precision", ["parity_tc", "fast"])
def test_adam_trajectory_at_c1_shape(precision):
    """10 Adam steps at the C1 MODEL shape (28x28, z = 2, 500x2, q 500x2; 32 images per step) against the oracle's
    trajectory from the same initial parameters and eps: PARITY_TC is inside the north-star 1e-4 on every parameter;
    FAST (one bf16 pass) is not expected to be (SURVEY 7.2 measured 3.6e-4 for it in emulation: Adam's m / sqrt(v)
    turns relative gradient noise on near-zero gradients into O(lr) steps), so it is bounded by what a 10-step
    trajectory at lr = 1e-4 allows -- > 99.9 % of parameters inside 1e-4, the largest deviation below 10 lr, the
    total update within 2 % of the reference's in L2 -- and its measured values are recorded."""
    import bench
    import json
    dev = _cuda()
    SF = _sf()
    c = dict(bench.CONFIGS["c1"])
    P, B, steps, lr = c["n"] * c["n"], 32, 10, 1e-4
    dec, enc = O.init_params(P, c["Z"] + 3, c["Z"], c["H"], c["L"], c["Hq"], c["Lq"], c["C"], seed=2)
    cfg = O.StepConfig(family="mnist", theta_prior=c["theta_prior"])
    grid = O.make_grid(c["n"], c["n"])
    ys = [bench.synth_images(c, B, torch.device("cpu"), 500 + t) for t in range(steps)]
    epss = [torch.randn(B, c["Z"] + 3, generator=torch.Generator().manual_seed(1000 + t)) for t in range(steps)]
    init = torch.cat([p.reshape(-1) for p in O.flatten_params(dec, enc)]).clone()
    dec_o, enc_o, elbos_ref = O.train_steps(cfg, dec, enc, grid, ys, epss, lr=lr)
    ref = torch.cat([p.reshape(-1) for p in O.flatten_params(dec_o, enc_o)])
    dd, ee = _to_dev(dec, enc, dev)
    params = dd.flat() + [t for pair in ee for t in pair]
    flat = torch.cat([p.reshape(-1) for p in params]).clone()
    gflat, m, v = torch.zeros_like(flat), torch.zeros_like(flat), torch.zeros_like(flat)
    def views(buf):
        out, off = [], 0
        for p in params:
            out.append(buf[off:off + p.numel()].view_as(p))
            off += p.numel()
        return out
    n_dec = len(dd.flat())
    pv, gv = views(flat), views(gflat)
    dd = SF.DecoderTensors.from_flat(pv[:n_dec], dd.latent_w is not None, len(dd.hidden))
    ee = [(pv[n_dec + i], pv[n_dec + i + 1]) for i in range(0, len(pv) - n_dec, 2)]
    gd = SF.DecoderTensors.from_flat(gv[:n_dec], dd.latent_w is not None, len(dd.hidden))
    ge = [(gv[n_dec + i], gv[n_dec + i + 1]) for i in range(0, len(gv) - n_dec, 2)]
    elbos = []
    for t in range(steps):
        stats, _, _ = SF.run_step(_spec(cfg, precision), dd, ee, grid.to(dev), ys[t].to(dev), epss[t].to(dev),
                                  grad_dec=gd, grad_enc=ge)
        SF.adam_step(flat, gflat, m, v, lr, t + 1)
        elbos.append(float(stats[:, 2].mean()))
    diff = (flat.cpu() - ref).abs()
    rec = {"shape": "c1", "precision": precision, "max_abs_dparam": float(diff.max()), "mean_abs_dparam": float(diff.mean()),
           "frac_within_1e-4": float((diff < 1e-4).float().mean()),
           "rel_l2_error_of_10_step_update": float((flat.cpu() - ref).norm() / (ref - init).norm()),
           "max_rel_elbo_err": float(np.max(np.abs(np.array(elbos) - np.array(elbos_ref)) / np.abs(np.array(elbos_ref))))}
    print(rec)
    try:
        out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
        os.makedirs(out_dir, exist_ok=True)
        with open(os.path.join(out_dir, f"parity_trajectory_c1_{precision}.json"), "w") as f:
            json.dump(rec, f)
    except OSError:
        pass
    assert rec["max_rel_elbo_err"] < (1e-3 if precision == "fast" else 2e-5), rec
    if precision == "parity_tc":
        assert rec["max_abs_dparam"] < 1e-4, rec
    else:
        assert rec["frac_within_1e-4"] > 0.999 and rec["max_abs_dparam"] < 1e-3 and \
            rec["rel_l2_error_of_10_step_update"] < 0.02, rec


# ---- full BASELINE sizes: size-independent properties + fast-vs-parity agreement -------------------------
def _full_case(cfgname):
    import bench
    c = dict(bench.CONFIGS[cfgname])
    P = c["n"] * c["n"]
    dec, enc = O.init_params(P * c["Cin"], c["Z"] + 3, c["Z"], c["H"], c["L"], c["Hq"], c["Lq"], c["C"], seed=1)
    y = bench.synth_images(c, c["B"], torch.device("cpu"), 1234)
    eps = torch.randn(c["B"], c["Z"] + 3, generator=torch.Generator().manual_seed(1000))
    cfg = O.StepConfig(family=c["family"], theta_prior=c["theta_prior"])
    return c, dec, enc, O.make_grid(c["n"], c["n"]), y, eps, cfg


def test_c2_full_size_properties():
    """configs[1] at its full size (28x28, z=100, 500x2, B=1024): the minibatch decomposes over images
    (what data parallelism relies on), fast and parity precision agree within the north-star tolerance, and
    the outputs respect their ranges."""
    c, dec, enc, grid, y, eps, cfg = _full_case("c2")
    B = c["B"]
    sf, yh, gf = _run_cuda(cfg, dec, enc, grid, y, eps, "fast")
    sp, _, gp = _run_cuda(cfg, dec, enc, grid, y, eps, "parity")
    rel = (sf[:, 2] - sp[:, 2]).abs() / sp[:, 2].abs()
    assert float(rel.max()) <= 1e-3, f"fast vs parity per-image ELBO: {float(rel.max()):.2e}"
    assert float(sf[:, 1].min()) >= 0.0                      # KL terms are non-negative
    assert float(yh.min()) > 0.0 and float(yh.max()) < 1.0   # sigmoid outputs
    # two half batches with grad_scale = 1/B: same per-image values, gradients add up
    h = B // 2
    sa, _, ga = _run_cuda(cfg, dec, enc, grid, y[:h], eps[:h], "fast", grad_scale=1.0 / B)
    sb, _, gb = _run_cuda(cfg, dec, enc, grid, y[h:], eps[h:], "fast", grad_scale=1.0 / B)
    np.testing.assert_allclose(torch.cat([sa, sb]).numpy(), sf.numpy(), rtol=2e-6, atol=1e-4)
    for i, (a, b, full) in enumerate(zip(ga, gb, gf)):
        scale = float(full.abs().max()) + 1e-12
        assert float((a + b - full).abs().max()) <= 2e-3 * scale, f"grad {i}"
    # a permutation of the batch permutes the per-image results
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(3))
    sq, _, _ = _run_cuda(cfg, dec, enc, grid, y[perm], eps[perm], "fast")
    np.testing.assert_allclose(sq.numpy(), sf[perm].numpy(), rtol=2e-6, atol=1e-4)
    # fast gradients track the fp32 ones
    for i, (a, b) in enumerate(zip(gf, gp)):
        scale = float(b.abs().max()) + 1e-12
        assert float((a - b).abs().max()) <= 5e-2 * scale, f"fast vs parity grad {i}"


@pytest.mark.parametrize("cfgname,B", [("c3", 12), ("c5", 8)])
def test_particle_configs_at_full_model_size_against_oracle(cfgname, B):
    """C3 (fit-noise) and C5 (39x39 CTF) model shapes, small batch: per-image ELBO vs the oracle."""
    c, dec, enc, grid, y, eps, cfg = _full_case(cfgname)
    y, eps = y[:B], eps[:B]
    kw = {}
    if c.get("ctf"):
        kw["ctf"] = 0.03 * torch.randn(B, 1, c["ctf"], c["ctf"], generator=torch.Generator().manual_seed(5))
    out, _ = O.step_grads(cfg, dec, enc, grid, y, eps, **kw)
    ref = (out["logp_i"] - out["kl_i"]).numpy()
    for precision, tol in (("parity", 2e-5), ("fast", 1e-3)):
        stats, _, _ = _run_cuda(cfg, dec, enc, grid, y, eps, precision, **kw)
        rel = np.abs(stats[:, 2].numpy() - ref) / np.abs(ref)
        assert rel.max() <= tol, f"{cfgname} {precision}: {rel.max():.2e}"

