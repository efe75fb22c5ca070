"""The library's kernel sources and host orchestration, executed on the CPU (tests/simt_emu: the SIMT .cu files compiled
with g++ behind a fiber-based shim, tc_gemm.cu on a host model of tcgen05 / TMA / mbarriers) and held to the same golden
fixtures and tolerances as the GPU parity tests.  This is how kernels written without GPU access are checked before they
first run on a B200; it covers indexing, reductions, barrier choreography and the call sequence of api.cu, not GPU-only
behaviour (memory-model races, occupancy, timing).
"""
import numpy as np
import pytest
import torch

from oracle import svae_oracle as O
from tests import emu_backend
from tests.helpers import cfg_of, golden_grads, load_case, option_cfg, oracle_params

CASES = [("mnist_rt", "mnist"), ("mnist_r", "mnist"), ("mnist_t", "mnist"), ("mnist_none", "mnist"),
         ("mnist_leaky_L3", "mnist"), ("particles_plain", "particles"), ("particles_fitnoise", "particles"),
         ("particles_ctf", "particles"), ("particles_mask", "particles"), ("particles_augment", "particles"),
         ("particles_zscale0", "particles"), ("particles_t_only", "particles"), ("particles_r_only", "particles"),
         ("galaxy_rgb", "galaxy")]
OPTION_CASES = ["particles_opt_resid", "particles_opt_expand", "particles_opt_bilinear", "particles_opt_softplus",
                "particles_opt_all"]


@pytest.fixture()
def emu(monkeypatch):
    return emu_backend.install(monkeypatch)


def _params(dec, enc):
    import spatial_vae.functional as SF
    d = SF.DecoderTensors(dec["coord_w"].clone(), dec["coord_b"].clone(),
                          dec["latent_w"].clone() if dec["latent_w"] is not None else None,
                          [(w.clone(), b.clone()) for w, b in dec["hidden"]], dec["out_w"].clone(), dec["out_b"].clone(),
                          dec["bilinear_w"].clone() if dec.get("bilinear_w") is not None else None)
    e = [(w.clone(), b.clone()) for w, b in enc]
    gd = SF.DecoderTensors.from_flat([torch.zeros_like(t) for t in d.flat()], *d.layout())
    ge = [(torch.zeros_like(w), torch.zeros_like(b)) for w, b in e]
    return d, e, gd, ge


def _run(cfg, dec, enc, grid, y, eps, precision, chunk=0, **kw):
    import spatial_vae.functional as SF
    from spatial_vae import _lib as L
    d, e, gd, ge = _params(dec, enc)
    spec = SF.StepSpec(family=cfg.family, rotate=cfg.rotate, translate=cfg.translate, dx_scale=cfg.dx_scale,
                       theta_prior=cfg.theta_prior, z_scale=cfg.z_scale, activation=L.ACT_CODES[cfg.activation],
                       softplus=cfg.softplus, precision=precision, chunk_images=chunk, resid=cfg.resid)
    stats, y_hat, _ = SF.run_step(spec, d, e, grid, y, eps, grad_dec=gd, grad_enc=ge, want_y_hat=True, **kw)
    return stats, y_hat, gd.flat() + [t for pr in ge for t in pr]


def _inputs(d):
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


def _check_parity(name, d, cfg, stats, y_hat, grads):
    for col, key in ((2, "elbo"), (0, "logp"), (1, "kl")):
        np.testing.assert_allclose(float(stats[:, col].mean()), float(d[key]), rtol=2e-5, atol=2e-6, err_msg=key)
    if "y_hat" in d:
        np.testing.assert_allclose(y_hat.reshape(d["y_hat"].shape).numpy(), d["y_hat"], rtol=1e-5, atol=1e-6)
    ref = golden_grads(d)
    assert len(grads) == len(ref)
    for i, (g, r) in enumerate(zip(grads, ref)):
        np.testing.assert_allclose(g.numpy(), r.numpy(), rtol=5e-4, atol=5e-6, err_msg=f"{name} grad {i}")


@pytest.mark.parametrize("name,family", CASES)
def test_emulated_step_matches_reference_golden_parity_precision(emu, name, family):
    d = load_case(name)
    dec, enc = oracle_params(d)
    cfg = cfg_of(d, family)
    grid, y, eps, kw = _inputs(d)
    stats, y_hat, grads = _run(cfg, dec, enc, grid, y, eps, "parity", **kw)
    _check_parity(name, d, cfg, stats, y_hat, grads)


@pytest.mark.parametrize("name", OPTION_CASES)
def test_emulated_option_step_matches_reference_golden(emu, name):
    d = load_case(name)
    dec, enc = oracle_params(d)
    cfg = option_cfg(d)
    grid, y, eps, kw = _inputs(d)
    stats, y_hat, grads = _run(cfg, dec, enc, grid, y, eps, "parity", **kw)
    _check_parity(name, d, cfg, stats, y_hat, grads)


@pytest.mark.parametrize("name,family", CASES + [(n, "particles") for n in OPTION_CASES])
def test_emulated_step_fast_precision_orchestration(emu, name, family):
    """FAST precision (bf16 activations, split3 encoder, fused output dot, bf16 SIMT kernels, the tcgen05 GEMM kernel on
    its host model): north-star tolerance on the ELBO, loose on the gradients."""
    d = load_case(name)
    dec, enc = oracle_params(d)
    cfg = option_cfg(d, family) if name in OPTION_CASES else cfg_of(d, family)
    grid, y, eps, kw = _inputs(d)
    stats, _, grads = _run(cfg, dec, enc, grid, y, eps, "fast", **kw)
    assert abs(float(stats[:, 2].mean()) - float(d["elbo"])) <= 1e-3 * abs(float(d["elbo"])) + 1e-4
    for i, (g, r) in enumerate(zip(grads, golden_grads(d))):
        scale = float(r.abs().max()) + 1e-6
        gtol = 3e-2 if cfg.activation == "tanh" else 1.5e-1
        assert float((g - r).abs().max()) <= gtol * scale, f"{name} grad {i}"


def _random_case(family, B, n, H, L, Z, Hq, C=1, seed=0):
    P = n * n
    Cin = 3 if family == "galaxy" else 1
    dec, enc = O.init_params(P * Cin, Z + 3, Z, H, L, Hq, 2, C, seed=seed)
    g = torch.Generator().manual_seed(1234 + seed)
    y = torch.rand(B, P, 3, generator=g) if family == "galaxy" else torch.randn(B, P, generator=g)
    return dec, enc, O.make_grid(n, n), y, torch.randn(B, Z + 3, generator=g)


@pytest.mark.parametrize("fast_kernel", [False, True])
@pytest.mark.parametrize("masked", [False, True])
def test_emulated_ctf_40x40_k39(monkeypatch, tmp_path, fast_kernel, masked):
    """C5 geometry (40x40 images, 39x39 kernels): the shared-memory correlation kernel and the register-tiled
    likelihood_ctf_k<39> (SVAE_CTF_FAST=1, not yet run on a GPU) against the oracle."""
    import math
    monkeypatch.setenv("SVAE_CTF_FAST", "1" if fast_kernel else "0")
    emu_backend.install(monkeypatch, fresh_copy_dir=tmp_path)     # the switch is read once per loaded library
    B, n = 2, 40
    dec, enc, grid, y, eps = _random_case("particles", B, n, 24, 2, 2, 16, seed=5)
    ctf = 0.03 * torch.randn(B, 1, 39, 39, generator=torch.Generator().manual_seed(9))
    mask = None
    if masked:
        yy, xx = np.ogrid[:n, :n]
        mask = torch.from_numpy(np.sqrt((n / 2 - yy) ** 2 + (n / 2 - xx) ** 2) < n / 2).view(-1)
    cfg = O.StepConfig(family="particles", theta_prior=math.pi)
    out, ograds = O.step_grads(cfg, dec, enc, grid, y, eps, ctf=ctf, mask=mask)
    stats, _, grads = _run(cfg, dec, enc, grid, y, eps, "parity", ctf=ctf, mask=mask)
    ref = (out["logp_i"] - out["kl_i"]).numpy()
    assert (np.abs(stats[:, 2].numpy() - ref) / np.abs(ref)).max() <= 2e-5
    for i, (gg, r) in enumerate(zip(grads, ograds)):
        scale = float(r.abs().max()) + 1e-8
        assert float((gg - r).abs().max()) <= 1e-3 * scale, f"grad {i}"


@pytest.mark.parametrize("name", ["particles_opt_all", "particles_opt_bilinear", "mnist_rt"])
def test_emulated_step_is_chunk_invariant(emu, name):
    d = load_case(name)
    dec, enc = oracle_params(d)
    cfg = option_cfg(d) if name in OPTION_CASES else cfg_of(d, "mnist")
    grid, y, eps, kw = _inputs(d)
    s1, y1, g1 = _run(cfg, dec, enc, grid, y, eps, "parity", chunk=0, **kw)
    s1, y1, g1 = s1.clone(), y1.clone(), [g.clone() for g in g1]
    s2, y2, g2 = _run(cfg, dec, enc, grid, y, eps, "parity", chunk=max(1, y.shape[0] // 3), **kw)
    np.testing.assert_allclose(s1.numpy(), s2.numpy(), rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(y1.numpy(), y2.numpy(), rtol=1e-6, atol=1e-7)
    for a, b in zip(g1, g2):
        np.testing.assert_allclose(a.numpy(), b.numpy(), rtol=1e-4, atol=1e-6 * max(1.0, float(b.abs().max())))


@pytest.mark.parametrize("resid,expand,bilinear", [(False, False, False), (True, False, False), (False, True, False),
                                                    (False, False, True), (True, True, True)])
@pytest.mark.parametrize("precision", ["parity", "fast"])
def test_emulated_modules_forward_backward(emu, resid, expand, bilinear, precision):
    """SpatialGenerator / InferenceNetwork modules (svae_decoder_* / svae_encoder_* entries, explicit coordinates,
    gradients w.r.t. parameters, x and z) against the oracle's autograd."""
    import contextlib
    import io
    import spatial_vae.models as M
    torch.manual_seed(5)
    B, n, H, Z, L = 3, 6, 24, 4, 3
    P = n * n
    with contextlib.redirect_stdout(io.StringIO()):
        p = M.SpatialGenerator(Z, H, n_out=2, num_layers=L, resid=resid, expand_coords=expand, bilinear=bilinear)
        q = M.InferenceNetwork(P, Z + 3, 20, num_layers=3, resid=resid)
    p.precision = precision
    dec = O.decoder_params_from_state({k: v.detach().clone() for k, v in p.state_dict().items()})
    enc = O.encoder_params_from_state({k: v.detach().clone() for k, v in q.state_dict().items()})
    x, z, yin = torch.rand(B, P, 2) * 2 - 1, torch.randn(B, Z), torch.randn(B, P)
    w_y, w_q = torch.randn(B, P, 2), torch.randn(B, 2 * (Z + 3))
    xr, zr = x.clone().requires_grad_(), z.clone().requires_grad_()
    leaves = [t.requires_grad_() for t in O.flatten_params(dec, enc)]
    dec_r, enc_r = O.unflatten_like(dec, enc, leaves)
    y_ref = O.decoder_forward(dec_r, xr, zr, "tanh")
    mu, ls = O.encoder_forward(enc_r, yin, "tanh", resid)
    ref = torch.autograd.grad((y_ref * w_y).sum() + (torch.cat([mu, ls], 1) * w_q).sum(), leaves + [xr, zr])

    xd, zd = x.clone().requires_grad_(), z.clone().requires_grad_()
    y_got = p(xd, zd)
    mu_g, ls_g = q(yin)
    ((y_got * w_y).sum() + (torch.cat([mu_g, ls_g], 1) * w_q).sum()).backward()
    got = [t.grad for t in list(p.parameters()) + list(q.parameters())] + [xd.grad, zd.grad]
    tol = dict(rtol=1e-4, atol=1e-5) if precision == "parity" else dict(rtol=5e-2, atol=5e-3)
    np.testing.assert_allclose(y_got.detach().numpy(), y_ref.detach().numpy(), **tol)
    np.testing.assert_allclose(mu_g.detach().numpy(), mu.detach().numpy(), rtol=1e-4, atol=1e-5)
    assert len(got) == len(ref)
    for i, (g, r) in enumerate(zip(got, ref)):
        lim = (2e-4 if precision == "parity" else 3e-2) * (float(r.abs().max()) + 1e-6)
        assert float((g - r).abs().max()) <= lim, f"grad {i}: {float((g - r).abs().max())} > {lim}"


@pytest.mark.parametrize("name", ["particles_opt_all", "particles_fitnoise"])
def test_emulated_trainer_trajectory(emu, name):
    """10 Adam steps through Trainer (flat buffers, svae_adam_tick / svae_adam_step_graph): parameters within
    1e-4 of the oracle trajectory (the north-star gate), with every option on and with none."""
    import contextlib
    import io
    import spatial_vae.functional as SF
    import spatial_vae.models as M
    from spatial_vae.trainer import Trainer
    d = load_case(name)
    dec, enc = oracle_params(d)
    cfg = option_cfg(d)
    grid, y, _, _ = _inputs(d)
    B, H, Z = y.shape[0], dec["coord_w"].shape[0], dec["latent_w"].shape[1]
    with contextlib.redirect_stdout(io.StringIO()):
        p = M.SpatialGenerator(Z, H, n_out=dec["out_w"].shape[0], num_layers=len(dec["hidden"]) + 1, resid=cfg.resid,
                               expand_coords=dec["coord_w"].shape[1] == 5, bilinear=dec.get("bilinear_w") is not None,
                               softplus=cfg.softplus)
        q = M.InferenceNetwork(enc[0][0].shape[1], enc[-1][0].shape[0] // 2, enc[0][0].shape[0],
                               num_layers=len(enc) - 1, resid=cfg.resid)
    for t, r in zip(list(p.parameters()) + list(q.parameters()), O.flatten_params(dec, enc)):
        t.data.copy_(r)
    spec = SF.StepSpec(family="particles", rotate=cfg.rotate, translate=cfg.translate, dx_scale=cfg.dx_scale,
                       theta_prior=cfg.theta_prior, z_scale=cfg.z_scale, softplus=cfg.softplus, precision="parity",
                       resid=cfg.resid)
    tr = Trainer(p, q, spec, lr=1e-3)
    I = enc[-1][0].shape[0] // 2
    eps_seq = [torch.randn(B, I, generator=torch.Generator().manual_seed(1000 + s)) for s in range(10)]
    for eps in eps_seq:
        tr.step(grid, y, eps=eps)
    dec_o, enc_o, _ = O.train_steps(cfg, dec, enc, grid, [y] * 10, eps_seq, lr=1e-3)
    for t, r in zip(list(p.parameters()) + list(q.parameters()), O.flatten_params(dec_o, enc_o)):
        np.testing.assert_allclose(t.detach().numpy(), r.numpy(), rtol=0, atol=1e-4)


def test_integration_md_ctypes_stub_works_as_written(emu):
    """The binding INTEGRATION.md tells a reference maintainer to add is executed as printed (only the library path
    is redirected to the host build): one svae_step through raw ctypes on a reference-shaped module pair, per-image
    ELBO against the golden fixture."""
    import os
    import re
    import contextlib
    import io
    import spatial_vae.models as M
    from tests.simt_emu.build import build
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "INTEGRATION.md")).read()
    code = re.search(r"```python\n(# spatial_vae/_b200\.py.*?)```", text, flags=re.S).group(1)
    assert 'C.CDLL("libsvae_b200.so")' in code
    code = code.replace('C.CDLL("libsvae_b200.so")', f"C.CDLL({build()!r})")
    code = code.replace("torch.cuda.current_stream().cuda_stream", "0")          # no CUDA stream on the host build
    ns = {}
    exec(compile(code, "INTEGRATION.md", "exec"), ns)

    d = load_case("mnist_rt")
    ps = {k[2:]: torch.from_numpy(v) for k, v in d.items() if k.startswith("p.")}
    qs = {k[2:]: torch.from_numpy(v) for k, v in d.items() if k.startswith("q.")}
    with contextlib.redirect_stdout(io.StringIO()):
        p = M.SpatialGenerator(ps["latent_linear.weight"].shape[1], ps["coord_linear.weight"].shape[0], num_layers=2)
        q = M.InferenceNetwork(qs["layers.0.weight"].shape[1], qs["layers.4.weight"].shape[0] // 2,
                               qs["layers.0.weight"].shape[0], num_layers=2)
    p.load_state_dict(ps)
    q.load_state_dict(qs)
    dec_params = [t for t in p.parameters()]
    gdec, genc = ns["Dec"](), ns["Enc"]()
    grads = [torch.zeros_like(t) for t in list(p.parameters()) + list(q.parameters())]
    gdec.coord_w, gdec.coord_b, gdec.latent_w = (g.data_ptr() for g in grads[:3])
    gdec.hidden_w[0], gdec.hidden_b[0] = grads[3].data_ptr(), grads[4].data_ptr()
    gdec.out_w, gdec.out_b = grads[5].data_ptr(), grads[6].data_ptr()
    for i in range(3):
        genc.w[i], genc.b[i] = grads[7 + 2 * i].data_ptr(), grads[8 + 2 * i].data_ptr()
    stats = ns["step"](p, q, torch.from_numpy(d["grid"]), torch.from_numpy(d["y"]), torch.from_numpy(d["eps"]),
                       (gdec, genc), rotate=True, translate=True, theta_prior=float(d["theta_prior"]),
                       dx_scale=float(d["dx_scale"]))
    np.testing.assert_allclose(float(stats[:, 2].mean()), float(d["elbo"]), rtol=1e-3)
    ref = golden_grads(d)
    for i, (g, r) in enumerate(zip(grads, ref)):
        assert float((g - r).abs().max()) <= 3e-2 * (float(r.abs().max()) + 1e-6), i
    assert len(dec_params) == 7


@pytest.mark.parametrize("name,sms", [("particles_opt_resid", "148"), ("particles_opt_all", "2")])
def test_emulated_resid_on_the_tensor_core_route(monkeypatch, tmp_path, name, sms):
    """SVAE_RESID_TC=1: ResidLinear layers on the tcgen05 GEMMs (tc_gemm RES: the layer input tile is streamed into
    the forward epilogue and added before the activation; dX sees a bf16 W + I copy; the encoder's 3-term GEMMs see
    fp32 W + I).  Runs the real kernel on the host model of tests/simt_emu/tc_emu.h, on a 148-SM and on a 2-SM
    emulated device (several tiles per persistent CTA pair), against the reference-generated fixtures."""
    monkeypatch.setenv("SVAE_RESID_TC", "1")
    monkeypatch.setenv("SVAE_EMU_SMS", sms)
    emu_backend.install(monkeypatch, fresh_copy_dir=tmp_path)     # the switches are read once per loaded library
    d = load_case(name)
    dec, enc = oracle_params(d)
    cfg = option_cfg(d)
    grid, y, eps, kw = _inputs(d)
    stats, _, grads = _run(cfg, dec, enc, grid, y, eps, "fast", **kw)
    assert abs(float(stats[:, 2].mean()) - float(d["elbo"])) <= 1e-3 * abs(float(d["elbo"])) + 1e-4
    for i, (g, r) in enumerate(zip(grads, golden_grads(d))):
        assert float((g - r).abs().max()) <= 5e-2 * (float(r.abs().max()) + 1e-6), f"{name} grad {i}"


@pytest.mark.parametrize("name,family", [("mnist_rt", "mnist"), ("particles_fitnoise", "particles"), ("galaxy_rgb", "galaxy")])
def test_emulated_fused_backward_on_a_two_sm_device(monkeypatch, tmp_path, name, family):
    """The fused backward kernels of tc_bwd.cu (top-layer dW GEMM that builds delta in shared memory; transposed dX GEMM
    that recomputes h_0 and reduces delta_0 per image) on 2 emulated SMs: ONE CTA pair walks every tile, so the ring
    wrap-around, the running per-image sums across tiles and image boundaries inside tiles are all exercised."""
    monkeypatch.setenv("SVAE_EMU_SMS", "2")
    emu_backend.install(monkeypatch, fresh_copy_dir=tmp_path)
    d = load_case(name)
    dec, enc = oracle_params(d)
    cfg = cfg_of(d, family)
    grid, y, eps, kw = _inputs(d)
    stats, _, grads = _run(cfg, dec, enc, grid, y, eps, "fast", **kw)
    assert abs(float(stats[:, 2].mean()) - float(d["elbo"])) <= 1e-3 * abs(float(d["elbo"])) + 1e-4
    for i, (g, r) in enumerate(zip(grads, golden_grads(d))):
        assert float((g - r).abs().max()) <= 3e-2 * (float(r.abs().max()) + 1e-6), f"{name} grad {i}"


