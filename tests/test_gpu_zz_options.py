"""GPU tests of code that has NEVER run on a GPU (file name sorts last on purpose; tests/test_gpu_zy_late.py holds
the late tests that only use kernels already validated on hardware).

(1) The decoder / encoder options --resid, --expand-coords, --bilinear (+ --softplus) against the reference-generated
fixtures tests/golden/particles_opt_*.npz and the oracle (reference models.py:13-21,65-67,74-75,99-102,114-121), the
particle driver without theta / dx, whole-module pickles written by the reference.  (2) svae_ctf_filter.  (3) The
seeded random configurations of tests/test_emu_fuzz.py on the device.  (4) ResidLinear on the tcgen05 route
(SVAE_RESID_TC=1).  Every body has passed on the emulation (tests/test_emu_gpu_suite.py) and on a B200 (round 1's driver run), after which
the host-side gate SVAE_UNVALIDATED_OPTIONS was removed.
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
from tests.helpers import golden_grads, load_case, option_cfg, oracle_params
from tests.test_gpu_api import _inject_normal, _nets, _script


OPTION_CASES = ["particles_opt_resid", "particles_opt_expand", "particles_opt_bilinear", "particles_opt_softplus",
                "particles_opt_all", "particles_t_only", "particles_r_only"]   # + the particle driver without theta / dx


def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def _dev_params(dec, enc, dev):
    import spatial_vae.functional as SF
    d = SF.DecoderTensors(dec["coord_w"].to(dev), dec["coord_b"].to(dev),
                          dec["latent_w"].to(dev) if dec["latent_w"] is not None else None,
                          [(w.to(dev), b.to(dev)) for w, b in dec["hidden"]], dec["out_w"].to(dev), dec["out_b"].to(dev),
                          dec["bilinear_w"].to(dev) if dec.get("bilinear_w") is not None else None)
    e = [(w.to(dev), b.to(dev)) for w, b in enc]
    gd = SF.DecoderTensors.from_flat([torch.zeros_like(t) for t in d.flat()], *d.layout())
    ge = [(torch.zeros_like(w), torch.zeros_like(b)) for w, b in e]
    return d, e, gd, ge


def _mask(d):
    return torch.from_numpy(d["mask"]) if "mask" in d else None


def _run(cfg, dec, enc, grid, y, eps, precision, chunk=0, mask=None):
    import spatial_vae.functional as SF
    from spatial_vae import _lib as L
    dev = _cuda()
    d, e, gd, ge = _dev_params(dec, enc, dev)
    spec = SF.StepSpec(family=cfg.family, rotate=cfg.rotate, translate=cfg.translate, dx_scale=cfg.dx_scale,
                       theta_prior=cfg.theta_prior, z_scale=cfg.z_scale, activation=L.ACT_CODES[cfg.activation],
                       softplus=cfg.softplus, precision=precision, chunk_images=chunk, resid=cfg.resid)
    stats, y_hat, _ = SF.run_step(spec, d, e, grid.to(dev), y.to(dev), eps.to(dev), grad_dec=gd, grad_enc=ge,
                                  want_y_hat=True, mask=mask.to(dev) if mask is not None else None)
    torch.cuda.synchronize()
    return stats.cpu(), y_hat.cpu(), [g.cpu() for g in gd.flat()] + [t.cpu() for pr in ge for t in pr]


def _inputs(d):
    t = lambda k: torch.from_numpy(d[k]).float()
    return t("grid"), t("y"), t("eps")


@pytest.mark.parametrize("name", OPTION_CASES)
def test_option_step_matches_reference_golden_parity_precision(name):
    d = load_case(name)
    dec, enc = oracle_params(d)
    cfg = option_cfg(d)
    grid, y, eps = _inputs(d)
    stats, _, grads = _run(cfg, dec, enc, grid, y, eps, "parity", mask=_mask(d))
    for col, key in ((2, "elbo"), (0, "logp"), (1, "kl")):
        np.testing.assert_allclose(float(stats[:, col].mean()), float(d[key]), rtol=2e-5, atol=2e-6, err_msg=key)
    ref = golden_grads(d)
    assert len(grads) == len(ref)
    for i, (g, r) in enumerate(zip(grads, ref)):
        np.testing.assert_allclose(g.numpy(), r.numpy(), rtol=5e-4, atol=5e-6, err_msg=f"{name} grad {i}")


@pytest.mark.parametrize("name", OPTION_CASES)
def test_option_step_matches_reference_golden_fast_precision(name):
    d = load_case(name)
    dec, enc = oracle_params(d)
    cfg = option_cfg(d)
    grid, y, eps = _inputs(d)
    stats, _, grads = _run(cfg, dec, enc, grid, y, eps, "fast", mask=_mask(d))
    assert abs(float(stats[:, 2].mean()) - float(d["elbo"])) <= 1e-3 * abs(float(d["elbo"])) + 1e-4
    for i, (g, r) in enumerate(zip(grads, golden_grads(d))):
        scale = float(r.abs().max()) + 1e-6
        assert float((g - r).abs().max()) <= 5e-2 * scale, f"{name} grad {i}"


@pytest.mark.parametrize("name", ["particles_opt_all", "particles_opt_bilinear"])
def test_option_step_is_chunk_invariant(name):
    """The per-image moments and per-image coordinate weights are indexed by the image's position in the CALL, not in
    the chunk: splitting the decoder pass must not change anything beyond summation order."""
    d = load_case(name)
    dec, enc = oracle_params(d)
    cfg = option_cfg(d)
    grid, y, eps = _inputs(d)
    s1, y1, g1 = _run(cfg, dec, enc, grid, y, eps, "parity", chunk=0)
    s2, y2, g2 = _run(cfg, dec, enc, grid, y, eps, "parity", chunk=max(1, y.shape[0] // 3))
    np.testing.assert_allclose(s1.numpy(), s2.numpy(), rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(y1.numpy(), y2.numpy(), rtol=1e-6, atol=1e-7)
    for a, b in zip(g1, g2):
        np.testing.assert_allclose(a.numpy(), b.numpy(), rtol=1e-4, atol=1e-6 * max(1.0, float(b.abs().max())))


@pytest.mark.parametrize("resid,expand,bilinear", [(True, False, False), (False, True, False), (False, False, True),
                                                    (True, True, True)])
@pytest.mark.parametrize("precision", ["parity", "fast"])
def test_option_modules_forward_backward_against_oracle(resid, expand, bilinear, precision):
    """SpatialGenerator / InferenceNetwork as modules (svae_decoder_forward/backward, svae_encoder_forward/backward)
    with explicit coordinates: outputs and every gradient (parameters, x, z) against the oracle's autograd."""
    import spatial_vae.models as M
    dev = _cuda()
    torch.manual_seed(5)
    B, n, H, Z, L = 3, 12, 96, 4, 3
    P = n * n
    with contextlib.redirect_stdout(io.StringIO()):
        p = M.SpatialGenerator(Z, H, n_out=2, num_layers=L, resid=resid, expand_coords=expand, bilinear=bilinear)
        q = M.InferenceNetwork(P, Z + 3, 80, num_layers=3, resid=resid)
    p.precision = precision
    dec = O.decoder_params_from_state({k: v.detach().clone() for k, v in p.state_dict().items()})
    enc = O.encoder_params_from_state({k: v.detach().clone() for k, v in q.state_dict().items()})
    x = (torch.rand(B, P, 2) * 2 - 1)
    z = torch.randn(B, Z)
    yin = torch.randn(B, P)
    w_y = torch.randn(B, P, 2)
    w_q = torch.randn(B, 2 * (Z + 3))

    # oracle
    xr, zr = x.clone().requires_grad_(), z.clone().requires_grad_()
    leaves = [t.requires_grad_() for t in O.flatten_params(dec, enc)]
    dec_r, enc_r = O.unflatten_like(dec, enc, leaves)
    y_ref = O.decoder_forward(dec_r, xr, zr, "tanh")
    mu, ls = O.encoder_forward(enc_r, yin, "tanh", resid)
    loss = (y_ref * w_y).sum() + (torch.cat([mu, ls], 1) * w_q).sum()
    ref = torch.autograd.grad(loss, leaves + [xr, zr])

    # library
    p, q = p.to(dev), q.to(dev)
    xd, zd = x.to(dev).requires_grad_(), z.to(dev).requires_grad_()
    y_got = p(xd, zd)
    mu_g, ls_g = q(yin.to(dev))
    loss_g = (y_got * w_y.to(dev)).sum() + (torch.cat([mu_g, ls_g], 1) * w_q.to(dev)).sum()
    loss_g.backward()
    torch.cuda.synchronize()
    got = [t.grad.cpu() for t in list(p.parameters()) + list(q.parameters())] + [xd.grad.cpu(), zd.grad.cpu()]

    tol = dict(rtol=1e-4, atol=1e-5) if precision == "parity" else dict(rtol=5e-2, atol=5e-3)
    np.testing.assert_allclose(y_got.detach().cpu().numpy(), y_ref.detach().numpy(), **tol)
    np.testing.assert_allclose(mu_g.detach().cpu().numpy(), mu.detach().numpy(), rtol=1e-4, atol=1e-5)
    assert len(got) == len(ref)
    for i, (g, r) in enumerate(zip(got, ref)):
        scale = float(r.abs().max()) + 1e-6
        lim = (1e-3 if precision == "parity" else 5e-2) * scale
        assert float((g - r).abs().max()) <= lim, f"grad {i}: {float((g - r).abs().max())} > {lim}"


def test_option_trainer_trajectory_matches_oracle():
    """10 Adam steps with every option on (Trainer, flat buffers): parameters within 1e-4 of the oracle's."""
    import spatial_vae.functional as SF
    import spatial_vae.models as M
    from spatial_vae.trainer import Trainer
    dev = _cuda()
    d = load_case("particles_opt_all")
    dec, enc = oracle_params(d)
    cfg = option_cfg(d)
    grid, y, _ = _inputs(d)
    B, H, Z = y.shape[0], dec["coord_w"].shape[0], dec["latent_w"].shape[1]
    with contextlib.redirect_stdout(io.StringIO()):
        p = M.SpatialGenerator(Z, H, n_out=dec["out_w"].shape[0], num_layers=len(dec["hidden"]) + 1, resid=cfg.resid,
                               expand_coords=dec["coord_w"].shape[1] == 5, bilinear=dec.get("bilinear_w") is not None,
                               softplus=cfg.softplus)
        q = M.InferenceNetwork(enc[0][0].shape[1], enc[-1][0].shape[0] // 2, enc[0][0].shape[0],
                               num_layers=len(enc) - 1, resid=cfg.resid)
    for t, r in zip(list(p.parameters()) + list(q.parameters()), O.flatten_params(dec, enc)):
        t.data.copy_(r)
    p, q = p.to(dev), q.to(dev)
    spec = SF.StepSpec(family="particles", rotate=cfg.rotate, translate=cfg.translate, dx_scale=cfg.dx_scale,
                       theta_prior=cfg.theta_prior, z_scale=cfg.z_scale, softplus=cfg.softplus, precision="parity",
                       resid=cfg.resid)
    tr = Trainer(p, q, spec, lr=1e-3)
    I = enc[-1][0].shape[0] // 2
    eps_seq = [torch.randn(B, I, generator=torch.Generator().manual_seed(1000 + s)) for s in range(10)]
    for eps in eps_seq:
        tr.step(grid.to(dev), y.to(dev), eps=eps.to(dev))
    torch.cuda.synchronize()
    dec_o, enc_o, _ = O.train_steps(cfg, dec, enc, grid, [y] * 10, eps_seq, lr=1e-3)
    for t, r in zip(list(p.parameters()) + list(q.parameters()), O.flatten_params(dec_o, enc_o)):
        np.testing.assert_allclose(t.detach().cpu().numpy(), r.numpy(), rtol=0, atol=1e-4)


def test_particles_command_line_with_every_option():
    """train_particles.py --resid --expand-coords --bilinear --softplus --fit-noise on synthetic data: two epochs,
    finite ELBO, and the training ELBO improves at a learning rate where six steps are enough to see it (the flags
    reach the networks, the trainer and the fused step)."""
    import importlib.util
    import math
    _cuda()
    pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "spatial-vae_b200")
    spec = importlib.util.spec_from_file_location("cli_train_particles_opt", os.path.join(pkg, "train_particles.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    argv = ["--synthetic", "96", "--synthetic-size", "12", "--num-epochs", "2", "--minibatch-size", "40",
            "--p-hidden-dim", "64", "--p-num-layers", "3", "--q-hidden-dim", "64", "--q-num-layers", "3", "--fit-noise",
            "--resid", "--expand-coords", "--bilinear", "--softplus", "--seed", "0", "-l", "0.003"]
    with contextlib.redirect_stdout(io.StringIO()) as buf:
        mod.main(argv)
    lines = [l for l in buf.getvalue().splitlines() if "\t" in l]
    rows = [l.split("\t") for l in lines[1:]]
    assert len(rows) == 4
    vals = [float(r[-3]) for r in rows]
    assert all(math.isfinite(v) for v in vals) and vals[2] > vals[0]


@pytest.mark.parametrize("name", ["ref_pickle_plain", "ref_pickle_options"])
def test_reference_written_pickles_evaluate_here(name):
    """Whole-module .sav files written BY THE REFERENCE (tests/golden/ref_pickle_*.sav, produced by
    tests/golden/make_golden.py with the reference's classes) unpickle into this package's classes and give the
    reference's outputs: a user's trained models move over without conversion."""
    import spatial_vae.models as M
    dev = _cuda()
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    d = np.load(os.path.join(gold, name + ".npz"))
    p = torch.load(os.path.join(gold, name + "_generator.sav"), weights_only=False)
    q = torch.load(os.path.join(gold, name + "_inference.sav"), weights_only=False)
    assert type(p) is M.SpatialGenerator and type(q) is M.InferenceNetwork
    assert p.resid == q.resid == (name == "ref_pickle_options")
    p, q = p.to(dev), q.to(dev)
    p.precision = "parity"
    t = lambda k: torch.from_numpy(d[k]).to(dev)
    with torch.no_grad():
        y_hat = p(t("x"), t("z"))
        z_mu, z_logstd = q(t("y"))
    np.testing.assert_allclose(y_hat.cpu().numpy(), d["y_hat"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(z_mu.cpu().numpy(), d["z_mu"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(z_logstd.cpu().numpy(), d["z_logstd"], rtol=1e-5, atol=1e-6)


def test_device_ctf_filter_matches_the_reference_kernels():
    """svae_ctf_filter (all particles in one launch, fp64 separable inverse DFT) against the kernels the REFERENCE's
    ctf_filter produced (tests/golden/ctf_kernels.npz) and against the host rewrite for odd, even and rectangular
    sizes and a pixel-size scale."""
    import pandas as pd
    import spatial_vae.ctf as C
    import spatial_vae.functional as SF
    dev = _cuda()
    d = load_case("ctf_kernels")
    table = pd.DataFrame({c: d[c] for c in C.CTF_COLUMNS})
    got = SF.ctf_filter(table, int(d["n"]), int(d["m"]), device=dev).cpu().numpy()
    assert got.shape == d["kernels"].shape and got.dtype == np.float32
    np.testing.assert_allclose(got, d["kernels"], rtol=1e-5, atol=1e-8)
    rng = np.random.default_rng(5)
    N = 7
    table = pd.DataFrame(dict(defocus=rng.uniform(0.5, 4, N), cs=rng.uniform(0.01, 2.7, N),
                              voltage=rng.choice([120.0, 200.0, 300.0], N), apix=rng.uniform(1.0, 3.0, N),
                              bfactor=rng.uniform(0, 200, N), ampcont=rng.uniform(5, 15, N), dfdiff=np.zeros(N),
                              dfang=rng.uniform(0, 180, N)))
    for n, m, scale in ((39, 39, 1.0), (12, 12, 1.0), (9, 14, 2.0), (1, 5, 1.0)):
        ref = C.ctf_filter(table, n, m, scale=scale)
        got = SF.ctf_filter(table, n, m, scale=scale, device=dev).cpu().numpy()
        np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-8, err_msg=f"{n}x{m} scale {scale}")
    assert SF.ctf_filter(table.iloc[:0], 5, 5, device=dev).shape == (0, 5, 5)


# (4) the seeded random configurations of tests/test_emu_fuzz.py, on the real device
@pytest.mark.parametrize("seed", range(0, 48, 2))
def test_random_step_configuration_on_the_device(seed):
    _run_fuzz_step(seed)


def _run_fuzz_step(seed):
    """tests/test_emu_fuzz.py::test_random_configuration_matches_oracle with the tensors on the device."""
    import spatial_vae.functional as SF
    from spatial_vae import _lib as L
    import tests.test_emu_fuzz as F
    dev = _cuda()
    c = F._draw(seed)
    dec, enc, y, eps, kw, I = F._params(c, seed)
    if I == 0:
        pytest.skip("no latent at all")
    grid = O.make_grid(c["n_rows"], c["n_cols"])
    cfg = O.StepConfig(family=c["family"], rotate=c["rotate"], translate=c["translate"], theta_prior=0.9, dx_scale=0.2,
                       z_scale=0.7 if c["family"] != "mnist" else 1.0, activation=c["act"], softplus=c["softplus"],
                       resid=c["resid"])
    out, ograds = O.step_grads(cfg, dec, enc, grid, y, eps, **kw)
    ref = (out["logp_i"] - out["kl_i"]).numpy()
    for precision in ("parity", "fast"):
        d, e, gd, ge = _dev_params(dec, enc, dev)
        spec = SF.StepSpec(family=c["family"], rotate=c["rotate"], translate=c["translate"], theta_prior=0.9,
                           dx_scale=0.2, z_scale=0.7, activation=L.ACT_CODES[c["act"]], softplus=c["softplus"],
                           precision=precision, chunk_images=c["chunk"], resid=c["resid"])
        kw_lib = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in kw.items()}
        if "ctf" in kw_lib:
            kw_lib["ctf"] = kw_lib["ctf"].reshape(c["B"], 5, 5)
        stats, _, _ = SF.run_step(spec, d, e, grid.to(dev), y.to(dev), eps.to(dev), grad_dec=gd, grad_enc=ge, **kw_lib)
        torch.cuda.synchronize()
        got = stats[:, 2].cpu().numpy()
        tol = 3e-5 if precision == "parity" else 1e-3          # the north star's per-image ELBO tolerance
        assert (np.abs(got - ref) / np.maximum(np.abs(ref), 1.0)).max() <= tol, (c, precision, got, ref)
        if precision == "parity":
            grads = [g.cpu() for g in gd.flat()] + [t.cpu() for pr in ge for t in pr]
            for i, (a, b) in enumerate(zip(grads, ograds)):
                lim = 2e-3 * float(b.abs().max()) + 2e-6
                assert float((a - b).abs().max()) <= lim, (c, i, float((a - b).abs().max()), lim)


def test_resid_on_the_tensor_core_route_in_a_subprocess():
    """SVAE_RESID_TC=1 (read once per process, hence the subprocess): the ResidLinear fixtures, the resid module tests
    and the option trainer trajectory again with the skip connection added in the tcgen05 forward epilogue
    (tc_gemm RES) instead of running the fp32 kernels."""
    import subprocess
    import sys
    _cuda()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, SVAE_RESID_TC="1")
    sel = ("(golden_fast_precision and (resid or opt_all)) or "
           "(modules_forward_backward and fast and (True-False-False or True-True-True))")
    r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_zz_options.py", "-m", "gpu", "-q", "-x",
                        "-p", "no:cacheprovider", "-k", sel], cwd=root, env=env, capture_output=True, text=True,
                       timeout=1200)
    tail = "\n".join((r.stdout + r.stderr).splitlines()[-20:])
    assert r.returncode == 0 and " passed" in r.stdout and " failed" not in r.stdout, tail
    assert "4 passed" in r.stdout, tail
