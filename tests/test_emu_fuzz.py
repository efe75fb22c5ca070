"""Seeded random configurations of the fused step (shapes, depths, flags, options) on tests/simt_emu against the
oracle: per-image ELBO and every gradient in PARITY precision, per-image ELBO in FAST.  Covers corners the fixtures
do not: L = 1 (no hidden layer), Z = 0 (no latent), odd hidden widths (padded fp32 rows), non-square pixel counts,
B = 1, C = 3 Bernoulli, every activation, masks, CTF kernels, augmentation offsets, chunking, and the decoder/encoder
options in combination."""
import math

import numpy as np
import pytest
import torch

from oracle import svae_oracle as O
from tests import emu_backend


@pytest.fixture()
def emu(monkeypatch):
    return emu_backend.install(monkeypatch)


def _draw(seed):
    r = np.random.default_rng(seed)
    family = ["mnist", "particles", "galaxy"][seed % 3]
    cfg = dict(family=family, B=int(r.integers(1, 6)), H=int(r.choice([7, 16, 33, 64, 70])), L=int(r.integers(1, 5)),
               Z=int(r.choice([0, 1, 3, 5])), Hq=int(r.choice([9, 16, 24])), Lq=int(r.integers(1, 4)),
               rotate=bool(r.integers(0, 2)), translate=bool(r.integers(0, 2)),
               act=str(r.choice(["tanh", "leakyrelu", "relu", "sigmoid"])), chunk=int(r.choice([0, 1, 2])),
               softplus=False, resid=bool(r.integers(0, 4) == 0), expand=bool(r.integers(0, 3) == 0),
               bilinear=bool(r.integers(0, 3) == 0), mask=False, C=1)
    if family == "galaxy":
        cfg.update(n_rows=4, n_cols=4, C=3)
    elif family == "particles":
        cfg.update(n_rows=6, n_cols=6, C=int(r.choice([1, 2])), softplus=bool(r.integers(0, 2)),
                   mask=bool(r.integers(0, 2)), ctf=bool(r.integers(0, 2)), augment=bool(r.integers(0, 2)))
        if cfg["ctf"]:
            cfg["C"] = 1                     # the reference cannot combine CTF with fit-noise
    else:
        cfg.update(n_rows=int(r.choice([4, 5])), n_cols=int(r.choice([5, 6])))       # non-square images
    if cfg["Z"] == 0:
        cfg["bilinear"] = False
    return cfg


def _params(c, seed):
    g = torch.Generator().manual_seed(500 + seed)
    rnd = lambda *s, sc=0.4: torch.randn(*s, generator=g) * sc
    F = 5 if c["expand"] else 2
    H, Z, P = c["H"], c["Z"], c["n_rows"] * c["n_cols"]
    Cin = 3 if c["family"] == "galaxy" else 1
    I = Z + int(c["rotate"]) + 2 * int(c["translate"])
    dec = {"coord_w": rnd(H, F), "coord_b": rnd(H, sc=0.1), "latent_w": rnd(H, Z) if Z > 0 else None,
           "hidden": [(rnd(H, H, sc=1.0 / math.sqrt(H)), rnd(H, sc=0.1)) for _ in range(c["L"] - 1)],
           "out_w": rnd(c["C"], H, sc=1.0 / math.sqrt(H)), "out_b": rnd(c["C"], sc=0.1)}
    if c["bilinear"]:
        dec["bilinear_w"] = rnd(H, F, Z, sc=0.2)
    if c["resid"]:
        dec["resid"] = True
    Hq = c["Hq"]
    enc = [(rnd(Hq, P * Cin, sc=1.0 / math.sqrt(P * Cin)), rnd(Hq, sc=0.1))]
    enc += [(rnd(Hq, Hq, sc=1.0 / math.sqrt(Hq)), rnd(Hq, sc=0.1)) for _ in range(c["Lq"] - 1)]
    enc += [(rnd(2 * I, Hq, sc=0.3 / math.sqrt(Hq)), rnd(2 * I, sc=0.05))]
    y = torch.rand(c["B"], P, 3, generator=g) if Cin == 3 else \
        (torch.rand(c["B"], P, generator=g) if c["family"] == "mnist" else torch.randn(c["B"], P, generator=g))
    eps = torch.randn(c["B"], I, generator=g)
    mask = (torch.rand(P, generator=g) > 0.3) if c["mask"] else None
    kw = {"mask": mask}
    if c.get("ctf"):
        kw["ctf"] = 0.1 * torch.randn(c["B"], 1, 5, 5, generator=g)
    if c.get("augment") and c["rotate"]:
        kw["theta_offset"] = torch.rand(c["B"], generator=g) * 6.28
        kw["y_enc"] = torch.randn(c["B"], P, generator=g)
    return dec, enc, y, eps, kw, I


@pytest.mark.parametrize("seed", range(48))
def test_random_configuration_matches_oracle(emu, seed):
    import spatial_vae.functional as SF
    from spatial_vae import _lib as L
    c = _draw(seed)
    dec, enc, y, eps, kw, I = _params(c, seed)
    if I == 0:
        pytest.skip("no latent at all: the reference cannot build this network either")
    grid = O.make_grid(c["n_rows"], c["n_cols"])
    cfg = O.StepConfig(family=c["family"], rotate=c["rotate"], translate=c["translate"], theta_prior=0.9, dx_scale=0.2,
                       z_scale=0.7 if c["family"] != "mnist" else 1.0, activation=c["act"], softplus=c["softplus"],
                       resid=c["resid"])
    out, ograds = O.step_grads(cfg, dec, enc, grid, y, eps, **kw)
    ref = (out["logp_i"] - out["kl_i"]).numpy()
    for precision in ("parity", "fast"):
        d = SF.DecoderTensors(dec["coord_w"].clone(), dec["coord_b"].clone(),
                              dec["latent_w"].clone() if dec["latent_w"] is not None else None,
                              [(w.clone(), b.clone()) for w, b in dec["hidden"]], dec["out_w"].clone(),
                              dec["out_b"].clone(), dec["bilinear_w"].clone() if c["bilinear"] else None)
        e = [(w.clone(), b.clone()) for w, b in enc]
        gd = SF.DecoderTensors.from_flat([torch.zeros_like(t) for t in d.flat()], *d.layout())
        ge = [(torch.zeros_like(w), torch.zeros_like(b)) for w, b in e]
        spec = SF.StepSpec(family=c["family"], rotate=c["rotate"], translate=c["translate"], theta_prior=0.9,
                           dx_scale=0.2, z_scale=0.7, activation=L.ACT_CODES[c["act"]], softplus=c["softplus"],
                           precision=precision, chunk_images=c["chunk"], resid=c["resid"])
        kw_lib = dict(kw)
        if "ctf" in kw_lib:
            kw_lib["ctf"] = kw_lib["ctf"].reshape(c["B"], 5, 5)
        stats, _, _ = SF.run_step(spec, d, e, grid, y, eps, grad_dec=gd, grad_enc=ge, **kw_lib)
        got = stats[:, 2].numpy()
        scale = np.maximum(np.abs(ref), 1.0)
        tol = 3e-5 if precision == "parity" else 1e-2
        assert (np.abs(got - ref) / scale).max() <= tol, (c, precision, got, ref)
        if precision == "parity":
            grads = gd.flat() + [t for pr in ge for t in pr]
            assert len(grads) == len(ograds)
            for i, (a, b) in enumerate(zip(grads, ograds)):
                lim = 2e-3 * float(b.abs().max()) + 2e-6
                assert float((a - b).abs().max()) <= lim, (c, i, float((a - b).abs().max()), lim)


@pytest.mark.parametrize("seed", range(24))
def test_random_module_configuration_matches_oracle(emu, seed):
    """SpatialGenerator / InferenceNetwork as autograd modules (svae_decoder_forward/backward with explicit
    coordinates, svae_encoder_forward/backward) for random shapes and options, PARITY precision: outputs and the
    gradients w.r.t. every parameter, x and z against the oracle's autograd."""
    import contextlib
    import io
    import torch.nn as nn
    import spatial_vae.models as M
    r = np.random.default_rng(1000 + seed)
    B, P = int(r.integers(1, 4)), int(r.choice([5, 12, 70]))          # P = 70 spans two 64-row blocks
    H, L, Z = int(r.choice([7, 16, 33, 64])), int(r.integers(1, 4)), int(r.choice([0, 1, 4]))
    C = int(r.integers(1, 4))
    act_name = str(r.choice(["tanh", "leakyrelu", "relu", "sigmoid"]))
    act = {"tanh": nn.Tanh, "leakyrelu": nn.LeakyReLU, "relu": nn.ReLU, "sigmoid": nn.Sigmoid}[act_name]
    resid, expand = bool(r.integers(0, 3) == 0), bool(r.integers(0, 2))
    bilinear = bool(r.integers(0, 2)) and Z > 0
    softplus = bool(r.integers(0, 2))
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        p = M.SpatialGenerator(Z, H, n_out=C, num_layers=L, activation=act, softplus=softplus, resid=resid,
                               expand_coords=expand, bilinear=bilinear)
        q = M.InferenceNetwork(P, Z + 2, int(r.choice([6, 20])), num_layers=int(r.integers(1, 4)), activation=act,
                               resid=resid)
    p.precision = "parity"
    dec = O.decoder_params_from_state({k: v.detach().clone() for k, v in p.state_dict().items()})
    enc = O.encoder_params_from_state({k: v.detach().clone() for k, v in q.state_dict().items()})
    x, yin = torch.rand(B, P, 2) * 2 - 1, torch.randn(B, P)
    z = torch.randn(B, Z) if Z > 0 else None
    w_y, w_q = torch.randn(B, P, C), torch.randn(B, 2 * (Z + 2))
    xr = x.clone().requires_grad_()
    zr = z.clone().requires_grad_() if Z > 0 else None
    leaves = [t.requires_grad_() for t in O.flatten_params(dec, enc)]
    dec_r, enc_r = O.unflatten_like(dec, enc, leaves)
    y_ref = O.decoder_forward(dec_r, xr, zr, act_name, softplus=softplus)
    mu, ls = O.encoder_forward(enc_r, yin, act_name, resid)
    wrt = leaves + [xr] + ([zr] if Z > 0 else [])
    ref = torch.autograd.grad((y_ref * w_y).sum() + (torch.cat([mu, ls], 1) * w_q).sum(), wrt)

    xd = x.clone().requires_grad_()
    zd = z.clone().requires_grad_() if Z > 0 else None
    y_got = p(xd, zd if Z > 0 else torch.zeros(B, 0))
    mu_g, ls_g = q(yin)
    ((y_got * w_y).sum() + (torch.cat([mu_g, ls_g], 1) * w_q).sum()).backward()
    got = [t.grad for t in list(p.parameters()) + list(q.parameters())] + [xd.grad] + ([zd.grad] if Z > 0 else [])
    cfg = dict(B=B, P=P, H=H, L=L, Z=Z, C=C, act=act_name, resid=resid, expand=expand, bilinear=bilinear)
    np.testing.assert_allclose(y_got.detach().numpy(), y_ref.detach().numpy(), rtol=1e-4, atol=1e-5, err_msg=str(cfg))
    np.testing.assert_allclose(mu_g.detach().numpy(), mu.detach().numpy(), rtol=1e-4, atol=1e-5, err_msg=str(cfg))
    assert len(got) == len(ref)
    for i, (g, rr) in enumerate(zip(got, ref)):
        assert g is not None, (cfg, i)
        lim = 1e-3 * float(rr.abs().max()) + 2e-6
        assert float((g - rr).abs().max()) <= lim, (cfg, i, float((g - rr).abs().max()), lim)


@pytest.mark.parametrize("sms", ["148", "3"])
def test_random_tcgen05_gemm_shapes_on_the_host_model(monkeypatch, tmp_path, sms):
    """svae_gemm_bf16 (tc_gemm.cu on the tcgen05 host model) for random shapes: ragged M below and above one pair
    tile, every multiple of 64 up to 384 for N and K, all three modes, on a 148-SM and on a 3-SM emulated device (odd
    number of SMs: one CTA pair only, so every tile goes through the same persistent pair)."""
    import spatial_vae.functional as SF
    monkeypatch.setenv("SVAE_EMU_SMS", sms)
    emu_backend.install(monkeypatch, fresh_copy_dir=tmp_path)
    r = np.random.default_rng(int(sms))
    g = torch.Generator().manual_seed(int(sms))
    for case in range(14):
        M = int(r.choice([1, 7, 127, 128, 129, 255, 256, 257, 300, 513]))
        N, K = int(r.choice([64, 128, 192, 320, 384])), int(r.choice([64, 128, 256, 384]))
        mode = case % 3
        if mode == 2:
            # dW: outf[M2, N2] += A[Kr, M2]^T B[Kr, N2] with ragged output sizes and a ragged reduction length
            Kr, M2, N2 = int(r.choice([5, 64, 200, 700])), int(r.choice([30, 64, 100])), int(r.choice([17, 64, 130]))
            lda, ldb = (M2 + 63) // 64 * 64, (N2 + 63) // 64 * 64
            A = torch.zeros(Kr, lda, dtype=torch.bfloat16)
            Bm = torch.zeros(Kr, ldb, dtype=torch.bfloat16)
            A[:, :M2] = (torch.randn(Kr, M2, generator=g) * 0.3).bfloat16()
            Bm[:, :N2] = (torch.randn(Kr, N2, generator=g) * 0.3).bfloat16()
            out = torch.randn(M2, N2, generator=g)
            ref = out + A[:, :M2].float().t() @ Bm[:, :N2].float()
            SF.gemm_bf16(2, A, Bm, M=M2, N=N2, K=Kr, out=out)
            np.testing.assert_allclose(out.numpy(), ref.numpy(), rtol=1e-4, atol=1e-4, err_msg=f"dw {Kr} {M2} {N2}")
            continue
        A = (torch.randn(M, K, generator=g) * 0.5).bfloat16()
        out = torch.full((M, N), float("nan"), dtype=torch.bfloat16)
        if mode == 0:
            W = (torch.randn(N, K, generator=g) / math.sqrt(K)).bfloat16()
            bias = torch.randn(N, generator=g)
            act = int(r.integers(0, 4))
            SF.gemm_bf16(0, A, W, M=M, N=N, K=K, bias=bias, activation=act, out=out)
            pre = A.float() @ W.float().t() + bias
            ref = [torch.tanh, lambda t: torch.where(t > 0, t, 0.01 * t), torch.relu, torch.sigmoid][act](pre)
        else:
            W = (torch.randn(K, N, generator=g) / math.sqrt(K)).bfloat16()
            aux = torch.tanh(torch.randn(M, N, generator=g)).bfloat16()
            SF.gemm_bf16(1, A, W, M=M, N=N, K=K, aux=aux, activation=0, out=out)
            ref = (A.float() @ W.float()) * (1 - aux.float() ** 2)
        assert torch.isfinite(out.float()).all(), (mode, M, N, K)
        np.testing.assert_allclose(out.float().numpy(), ref.numpy(), rtol=2e-2, atol=2e-2, err_msg=f"{mode} {M} {N} {K}")
