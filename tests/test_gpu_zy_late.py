"""GPU tests written after round 1's GPU budget was spent that only use kernels ALREADY validated on a B200 (the
default step with other flags / callers): display and generation helpers against reference outputs, train_epoch /
eval_model with the reference's signatures, softplus, ReLU / sigmoid activations.  Their bodies have passed on the
emulation (tests/test_emu_gpu_suite.py).  Sorts after the validated suites and before tests/test_gpu_zz_options.py
(kernels that have never run on a GPU), so a failure there cannot hide these.
"""
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import svae_oracle as O
from tests.helpers import golden_grads, load_case, oracle_params
from tests.test_gpu_api import _inject_normal, _nets, _script


def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def _dev():
    return _cuda()


def test_display_and_generation_helpers_match_the_reference():
    """minibatch_for_display / random_minibatch_generator of the galaxy driver against what the reference returned
    for the same injected normal draws (train_galaxy.py:131-183; q_net comes BEFORE p_net there, unlike
    train_mnist.py:93)."""
    import inspect
    dev = _dev()
    tg, tm = _script("train_galaxy"), _script("train_mnist")
    d = load_case("galaxy_rgb")
    p, q = _nets(d, dev)
    x, y = torch.from_numpy(d["grid"]).to(dev), torch.from_numpy(d["y"]).to(dev)
    with _inject_normal(torch.from_numpy(d["eps"])):
        disp = tg.minibatch_for_display(x, y, q, p, rotate=True, translate=True, z_scale=0.8)
    with _inject_normal(torch.from_numpy(d["z_rand"])):
        gen = tg.random_minibatch_generator(x, y, p, 4, z_scale=0.8)
    assert disp.shape == y.shape and gen.shape == y.shape
    np.testing.assert_allclose(disp.cpu().numpy(), d["display"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(gen.cpu().numpy(), d["generated"], rtol=1e-5, atol=1e-6)
    assert list(inspect.signature(tm.minibatch_for_display).parameters)[:4] == ["x", "y", "p_net", "q_net"]
    assert list(inspect.signature(tg.minibatch_for_display).parameters)[:4] == ["x", "y", "q_net", "p_net"]


def test_reference_style_epoch_loops_with_a_torch_optimizer(tmp_path):
    """train_epoch / eval_model with the reference's signatures: the caller's DataLoader and torch.optim.Adam, each
    minibatch through eval_minibatch + loss.backward() + optim.step() (train_mnist.py:127-226).  One epoch of three
    minibatches equals three oracle train steps; the returned running means are the reference's."""
    dev = _dev()
    tm = _script("train_mnist")
    d = load_case("mnist_rt")
    p, q = _nets(d, dev)
    dec, enc = oracle_params(d)
    grid = torch.from_numpy(d["grid"])
    g = torch.Generator().manual_seed(8)
    B, P = d["y"].shape
    ys = [(torch.rand(B, P, generator=g) > 0.7).float() * torch.rand(B, P, generator=g) for _ in range(3)]
    eps = torch.from_numpy(d["eps"])
    loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(torch.cat(ys).to(dev)), batch_size=B)
    optim = torch.optim.Adam(list(p.parameters()) + list(q.parameters()), lr=1e-3)
    cfg = O.StepConfig(family="mnist", theta_prior=float(d["theta_prior"]), dx_scale=float(d["dx_scale"]))
    with _inject_normal(eps):
        got = tm.train_epoch(loader, grid.to(dev), p, q, optim, rotate=True, translate=True,
                             dx_scale=float(d["dx_scale"]), theta_prior=float(d["theta_prior"]), epoch=0, num_epochs=1,
                             N=3 * B)
    # oracle: the same three steps; running means of per-step (elbo, -logp, kl)
    steps, dec_o, enc_o = [], dec, enc
    adam = O.AdamState(lr=1e-3)
    for y in ys:
        out, grads = O.step_grads(cfg, dec_o, enc_o, grid, y, eps)
        steps.append((float(out["elbo"]), -float(out["logp"]), float(out["kl"])))
        dec_o, enc_o = O.unflatten_like(dec_o, enc_o, adam.update(O.flatten_params(dec_o, enc_o), grads))
    np.testing.assert_allclose(got, np.mean(steps, axis=0), rtol=2e-4)
    for t, r in zip(list(p.parameters()) + list(q.parameters()), O.flatten_params(dec_o, enc_o)):
        np.testing.assert_allclose(t.detach().cpu().numpy(), r.numpy(), rtol=0, atol=1e-4)
    assert p.training and q.training
    with _inject_normal(eps):
        val = tm.eval_model(loader, grid.to(dev), p, q, rotate=True, translate=True, dx_scale=float(d["dx_scale"]),
                            theta_prior=float(d["theta_prior"]))
    ref = [O.step_forward(cfg, dec_o, enc_o, grid, y, eps) for y in ys]
    np.testing.assert_allclose(val[0], np.mean([float(r["elbo"]) for r in ref]), rtol=2e-4)
    assert not p.training and not q.training

    # particles: a loader of (y, ctf) pairs; galaxy: RGB
    tp, tg = _script("train_particles"), _script("train_galaxy")
    dp = load_case("particles_ctf")
    pp, qp = _nets(dp, dev)
    yp, ctf = torch.from_numpy(dp["y"]).to(dev), torch.from_numpy(dp["ctf"]).to(dev)
    lp = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(yp, ctf), batch_size=2)
    op = torch.optim.Adam(list(pp.parameters()) + list(qp.parameters()), lr=1e-3)
    gp = torch.from_numpy(dp["grid"]).to(dev)
    r1 = tp.train_epoch(lp, gp, None, pp, qp, op, N=yp.shape[0])
    r2 = tp.eval_model(lp, gp, None, pp, qp)
    dg = load_case("galaxy_rgb")
    pg, qg = _nets(dg, dev)
    lg = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(torch.from_numpy(dg["y"]).to(dev)), batch_size=2)
    og = torch.optim.Adam(list(pg.parameters()) + list(qg.parameters()), lr=1e-3)
    gg = torch.from_numpy(dg["grid"]).to(dev)
    r3 = tg.train_epoch(lg, gg, pg, qg, og, train_images_len=3)
    os.makedirs(tmp_path / "images")
    r4 = tg.eval_model(lg, gg, pg, qg, 4, to_save_image_samples=True, image_dims=[4, 4], epoch="01",
                       output_dir=str(tmp_path), save_label="t")
    assert all(math.isfinite(v) for r in (r1, r2, r3, r4) for v in r)


from tests.test_gpu_parity import _golden_inputs as _golden_inputs_parity, _random_case as _random_case_parity  # noqa: E402


def _run_parity_suite(*a, **k):
    import tests.test_gpu_parity as TP
    TP._cuda = _cuda            # same device helper (and the same emulation patch) as this module
    return TP._run_cuda(*a, **k)


# (3) softplus and the ReLU / sigmoid activations through the fused step
@pytest.mark.parametrize("precision", ["parity", "fast"])
def test_softplus_output_channel_matches_oracle(precision):
    """--softplus (models.py:129-130): softplus on output channel 0 AFTER the sigmoid, with and without fit-noise."""
    for C, seed in ((1, 11), (2, 12)):
        dec, enc, grid, y, eps = _random_case_parity("particles", 5, 10, 64, 2, 3, 32, C=C, seed=seed)
        cfg = O.StepConfig(family="particles", theta_prior=math.pi, softplus=True)
        out, ograds = O.step_grads(cfg, dec, enc, grid, y, eps)
        stats, _, grads = _run_parity_suite(cfg, dec, enc, grid, y, eps, precision)
        ref = (out["logp_i"] - out["kl_i"]).numpy()
        tol = 2e-5 if precision == "parity" else 1e-3
        assert (np.abs(stats[:, 2].numpy() - ref) / np.abs(ref)).max() <= tol
        gtol = 1e-3 if precision == "parity" else 5e-2
        for i, (g, r) in enumerate(zip(grads, ograds)):
            scale = float(r.abs().max()) + 1e-8
            assert float((g - r).abs().max()) <= gtol * scale, f"C={C} grad {i}"


@pytest.mark.parametrize("precision", ["parity", "fast"])
def test_softplus_golden_fixture(precision):
    """--softplus against the reference-generated fixture (3-layer decoder and encoder)."""
    from tests.helpers import option_cfg
    d = load_case("particles_opt_softplus")
    dec, enc = oracle_params(d)
    cfg = option_cfg(d)
    grid, y, eps, kw = _golden_inputs_parity(d)
    stats, _, grads = _run_parity_suite(cfg, dec, enc, grid, y, eps, precision)
    tol = 2e-5 if precision == "parity" else 1e-3
    assert abs(float(stats[:, 2].mean()) - float(d["elbo"])) <= tol * abs(float(d["elbo"])) + 1e-5
    if precision == "parity":
        for i, (g, r) in enumerate(zip(grads, golden_grads(d))):
            np.testing.assert_allclose(g.numpy(), r.numpy(), rtol=5e-4, atol=5e-6, err_msg=f"grad {i}")


def test_activation_variants_match_oracle():
    """ReLU and sigmoid hidden activations (train_galaxy.py:426-434) in both precisions."""
    for act in ("relu", "sigmoid"):
        dec, enc, grid, y, eps = _random_case_parity("galaxy", 4, 8, 96, 3, 3, 40, C=3, seed=21)
        cfg = O.StepConfig(family="galaxy", theta_prior=math.pi, activation=act)
        out, _ = O.step_grads(cfg, dec, enc, grid, y, eps)
        ref = (out["logp_i"] - out["kl_i"]).numpy()
        for precision, tol in (("parity", 2e-5), ("fast", 2e-3)):
            stats, _, _ = _run_parity_suite(cfg, dec, enc, grid, y, eps, precision)
            rel = (np.abs(stats[:, 2].numpy() - ref) / np.abs(ref)).max()
            assert rel <= tol, f"{act} {precision}: {rel:.2e}"
