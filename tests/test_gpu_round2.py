"""GPU parity tests added in round 2: the minibatch fetch bit for bit, the BASELINE model shapes
against the oracle (C2 z = 100, C4 galaxy 64x64x3 / 1000x4 / q 5000x2), the tensor-core parity mode, the in-kernel
eps generator, ResidLinear.forward and the --vanilla command lines.  All calls go through the C ABI."""
import contextlib
import io
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import svae_oracle as O
from tests.test_gpu_api import _script


def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


_dev = _cuda


def _sf():
    import spatial_vae.functional as SF
    return SF


def _dev_params(dec, enc, dev):
    SF = _sf()
    d = SF.DecoderTensors(dec["coord_w"].to(dev), dec["coord_b"].to(dev),
                          dec["latent_w"].to(dev) if dec["latent_w"] is not None else None,
                          [(w.to(dev), b.to(dev)) for w, b in dec["hidden"]], dec["out_w"].to(dev), dec["out_b"].to(dev))
    e = [(w.to(dev), b.to(dev)) for w, b in enc]
    gd = SF.DecoderTensors.from_flat([torch.zeros_like(t) for t in d.flat()], *d.layout())
    ge = [(torch.zeros_like(w), torch.zeros_like(b)) for w, b in e]
    return d, e, gd, ge


# ---- a11: minibatch fetch (reference train_mnist.py:334,395-396: DataLoader over a TensorDataset) --------------------
@pytest.mark.parametrize("n_src,row_len,n_idx", [(8192, 784, 1024), (300, 1521, 77), (64, 12288, 7), (50, 784, 0),
                                                 (10, 3, 10)])
def test_gather_rows_is_bit_exact(n_src, row_len, n_idx):
    """svae_gather_rows == src[index] bit for bit: images (784), 39x39 CTF rows (1521), RGB galaxy rows (12288), a
    ragged last minibatch, an empty slice (a rank without images), repeated and reversed indices."""
    dev = _cuda()
    SF = _sf()
    g = torch.Generator().manual_seed(n_src + row_len)
    src = torch.randn(n_src, row_len, generator=g).to(dev)
    idx = torch.randint(0, n_src, (n_idx,), generator=g)
    if n_idx >= 4:
        idx[:2] = idx[2:4]                      # duplicates
        idx = torch.cat([idx[: n_idx // 2], idx[: n_idx - n_idx // 2].flip(0)])
    got = SF.gather_rows(src, idx.to(dev))
    torch.cuda.synchronize()
    assert got.shape == (n_idx, row_len)
    assert torch.equal(got.cpu(), src.cpu()[idx])
    # a 3-D dataset (galaxy: N x P x 3) is fetched row-wise as well
    if row_len % 3 == 0 and n_idx > 0:
        src3 = src.view(n_src, row_len // 3, 3)
        got3 = SF.gather_rows(src3, idx.to(dev))
        assert torch.equal(got3.cpu().reshape(n_idx, -1), src.cpu()[idx])


# ---- BASELINE model shapes against the oracle --------------------------------------------------------------------------
def _config_case(name, B, seed=1):
    import bench
    c = dict(bench.CONFIGS[name])
    P = c["n"] * c["n"]
    dec, enc = O.init_params(P * c["Cin"], c["Z"] + 3, c["Z"], c["H"], c["L"], c["Hq"], c["Lq"], c["C"], seed=seed)
    y = bench.synth_images(c, B, torch.device("cpu"), 1234)
    eps = torch.randn(B, c["Z"] + 3, generator=torch.Generator().manual_seed(1000))
    cfg = O.StepConfig(family=c["family"], theta_prior=c["theta_prior"])
    return c, cfg, dec, enc, O.make_grid(c["n"], c["n"]), y, eps


@pytest.mark.parametrize("name,B", [("c2", 48), ("c4", 2)])
@pytest.mark.parametrize("precision,tol,gtol", [("fast", 1e-3, 0.05), ("parity_tc", 2e-5, 2e-3)])
def test_baseline_model_shapes_match_the_oracle(name, B, precision, tol, gtol):
    """C2 (z = 100, 500x2, 28x28) and C4 (galaxy 64x64x3, 1000x4, q 5000x2) at their MODEL shapes, a few images:
    per-image ELBO against the CPU oracle within the north-star tolerance (1e-3 relative) for the benchmarked FAST
    mode and within 2e-5 for PARITY_TC; gradients relative to each tensor's largest entry."""
    dev = _cuda()
    SF = _sf()
    from spatial_vae import _lib as L
    c, cfg, dec, enc, grid, y, eps = _config_case(name, B)
    out, ograds = O.step_grads(cfg, dec, enc, grid, y, eps)
    ref = (out["logp_i"] - out["kl_i"]).numpy()
    d, e, gd, ge = _dev_params(dec, enc, dev)
    spec = SF.StepSpec(family=c["family"], theta_prior=c["theta_prior"], precision=precision, activation=L.ACT_TANH)
    stats, _, _ = SF.run_step(spec, d, e, grid.to(dev), y.to(dev), eps.to(dev), grad_dec=gd, grad_enc=ge)
    torch.cuda.synchronize()
    got = stats[:, 2].cpu().numpy()
    rel = np.abs(got - ref) / np.abs(ref)
    assert rel.max() <= tol, (name, precision, rel.max())
    grads = [g.cpu() for g in gd.flat()] + [t.cpu() for pr in ge for t in pr]
    for i, (a, b) in enumerate(zip(grads, ograds)):
        scale = max(float(b.abs().max()), 1e-12)
        assert float((a - b).abs().max()) <= gtol * scale, (name, precision, i, float((a - b).abs().max()) / scale)


@pytest.mark.parametrize("precision,tol,gtol", [("fast", 1e-3, 0.05), ("parity_tc", 2e-5, 2e-3)])
def test_encoder_gemm_epilogues_small_and_large_batch(precision, tol, gtol):
    """The inference network's tensor-core GEMMs (models.py:46-54) finish a layer in two ways: with few output tiles
    (every benchmark minibatch) K is split over the idle CTA pairs and bias + activation are applied by the next
    kernel; with more tiles than half the GPU the GEMM's own epilogue does it.  A tiny model at 64 and at 9 600 images
    takes one path each; both against the CPU oracle."""
    dev = _cuda()
    SF = _sf()
    from spatial_vae import _lib as L
    n, H, Hq, Z = 4, 32, 64, 2
    P = n * n
    for B in (64, 9600):
        dec, enc = O.init_params(P, Z + 3, Z, H, 2, Hq, 2, 1, seed=5)
        g = torch.Generator().manual_seed(B)
        y = torch.rand(B, P, generator=g)
        eps = torch.randn(B, Z + 3, generator=g)
        cfg = O.StepConfig(family="mnist", theta_prior=math.pi / 4)
        grid = O.make_grid(n, n)
        out, ograds = O.step_grads(cfg, dec, enc, grid, y, eps)
        ref = (out["logp_i"] - out["kl_i"]).numpy()
        d, e, gd, ge = _dev_params(dec, enc, dev)
        spec = SF.StepSpec(family="mnist", theta_prior=math.pi / 4, precision=precision, activation=L.ACT_TANH)
        stats, _, _ = SF.run_step(spec, d, e, grid.to(dev), y.to(dev), eps.to(dev), grad_dec=gd, grad_enc=ge)
        torch.cuda.synchronize()
        rel = np.abs(stats[:, 2].cpu().numpy() - ref) / np.abs(ref)
        assert rel.max() <= tol, (B, precision, rel.max())
        grads = [t.cpu() for t in gd.flat()] + [t.cpu() for pr in ge for t in pr]
        for i, (a, b) in enumerate(zip(grads, ograds)):
            scale = max(float(b.abs().max()), 1e-12)
            assert float((a - b).abs().max()) <= gtol * scale, (B, precision, i, float((a - b).abs().max()) / scale)


def test_parity_tc_matches_parity_on_a_golden_step():
    """The tensor-core parity mode against the reference-written fixture, at the PARITY tolerances."""
    from tests.helpers import cfg_of, golden_grads, load_case, oracle_params
    dev = _cuda()
    SF = _sf()
    from spatial_vae import _lib as L
    for name, family in (("mnist_rt", "mnist"), ("particles_fitnoise", "particles"), ("galaxy_rgb", "galaxy"),
                         ("mnist_leaky_L3", "mnist")):
        d = load_case(name)
        dec, enc = oracle_params(d)
        cfg = cfg_of(d, family)
        dd, ee, gd, ge = _dev_params(dec, enc, dev)
        spec = SF.StepSpec(family=family, rotate=cfg.rotate, translate=cfg.translate, dx_scale=cfg.dx_scale,
                           theta_prior=cfg.theta_prior, z_scale=cfg.z_scale, activation=L.ACT_CODES[cfg.activation],
                           precision="parity_tc")
        t = lambda k: torch.from_numpy(d[k]).float().to(dev)
        stats, _, _ = SF.run_step(spec, dd, ee, t("grid"), t("y"), t("eps"), grad_dec=gd, grad_enc=ge)
        torch.cuda.synchronize()
        np.testing.assert_allclose(float(stats[:, 2].mean()), float(d["elbo"]), rtol=2e-5, atol=2e-6, err_msg=name)
        grads = [g.cpu() for g in gd.flat()] + [x.cpu() for pr in ge for x in pr]
        for i, (g, r) in enumerate(zip(grads, golden_grads(d))):
            np.testing.assert_allclose(g.numpy(), r.numpy(), rtol=2e-3, atol=2e-5 * max(1.0, float(r.abs().max())),
                                       err_msg=f"{name} grad {i}")


# ---- eps drawn in the kernel (reference train_mnist.py:38: eps ~ N(0, 1) per step) -----------------------------------
def test_in_kernel_eps_is_standard_normal_and_split_invariant():
    dev = _cuda()
    SF = _sf()
    c, cfg, dec, enc, grid, y, _ = _config_case("c2", 256)
    d, e, _, _ = _dev_params(dec, enc, dev)
    spec = SF.StepSpec(family="mnist", theta_prior=c["theta_prior"], precision="fast")
    # eps is recovered from the sampled latent further down: eps = (lat - mu) / sigma
    step = torch.zeros(1, dtype=torch.int32, device=dev)

    def latents(lo, hi, seed, step_value):
        step.fill_(step_value)
        _, _, lat = SF.run_step(spec, d, e, grid.to(dev), y[lo:hi].to(dev), None, want_latent=True, rng=(seed, step, lo))
        torch.cuda.synchronize()
        return lat.cpu()

    full = latents(0, 256, 1234, 0)
    # deterministic in (seed, step, image), different across steps and seeds
    # (the encoder head accumulates its split-K partial sums with atomics, so mu / sigma differ in the last bits)
    np.testing.assert_allclose(full.numpy(), latents(0, 256, 1234, 0).numpy(), rtol=1e-5, atol=1e-6)
    assert float((full - latents(0, 256, 1234, 1)).abs().mean()) > 0.1
    assert float((full - latents(0, 256, 99, 0)).abs().mean()) > 0.1
    # the same minibatch split 100 + 156 across two "ranks" draws the same numbers
    parts = torch.cat([latents(0, 100, 1234, 0), latents(100, 256, 1234, 0)])
    np.testing.assert_allclose(parts.numpy(), full.numpy(), rtol=1e-5, atol=1e-6)
    # moments of eps recovered through the oracle's encoder (lat = sigma * eps + mu)
    mu, ls = O.encoder_forward(enc, y)
    eps = ((full - mu) / torch.exp(ls)).numpy().ravel()
    n = eps.size
    assert abs(eps.mean()) < 5 / math.sqrt(n) and abs(eps.var() - 1) < 5 * math.sqrt(2 / n)
    assert abs((eps ** 3).mean()) < 0.1 and abs((eps ** 4).mean() - 3) < 0.2


# ---- the hook for an early decoder-gradient exchange (reference order train_mnist.py:147) ----------------------
def test_decoder_grads_event_marks_the_point_where_the_decoder_gradients_are_final():
    """A second stream that waits for SvaeStepInputs.decoder_grads_event and snapshots the decoder's gradient tensors
    sees their FINAL values (bit for bit), while the encoder backward is still to run on the main stream."""
    dev = _cuda()
    SF = _sf()
    c, cfg, dec, enc, grid, y, eps = _config_case("c2", 64)
    d, e, gd, ge = _dev_params(dec, enc, dev)
    spec = SF.StepSpec(family="mnist", theta_prior=c["theta_prior"], precision="fast")
    ev = torch.cuda.Event()
    ev.record()
    side = torch.cuda.Stream()
    torch.cuda.synchronize()
    SF.run_step(spec, d, e, grid.to(dev), y.to(dev), eps.to(dev), grad_dec=gd, grad_enc=ge, decoder_grads_event=ev)
    side.wait_event(ev)
    with torch.cuda.stream(side):
        snap = [t.clone() for t in gd.flat()]
    torch.cuda.synchronize()
    assert ev.query()
    for got, final in zip(snap, gd.flat()):
        assert torch.equal(got, final)
        assert torch.isfinite(final).all()
    assert any(float(t.abs().max()) > 0 for t in gd.flat())
    assert all(float(w.abs().max()) > 0 for w, _ in ge)
    # an event that was never recorded may have no CUDA handle yet (torch creates it lazily): refused, not ignored
    fresh = torch.cuda.Event()
    if not fresh.cuda_event:
        with pytest.raises(ValueError):
            SF.run_step(spec, d, e, grid.to(dev), y.to(dev), eps.to(dev), grad_dec=gd, grad_enc=ge,
                        decoder_grads_event=fresh)


# ---- ResidLinear.forward (reference models.py:13-21) ---------------------------------------------------------------------
def test_resid_linear_module_forward_and_backward():
    import spatial_vae.models as M
    import torch.nn as nn
    dev = _cuda()
    for act, ref_act in ((nn.Tanh, torch.tanh), (nn.LeakyReLU, nn.functional.leaky_relu)):
        torch.manual_seed(3)
        m = M.ResidLinear(48, 48, activation=act).to(dev)
        x = torch.randn(37, 48, device=dev, requires_grad=True)
        y = m(x)
        (y ** 2).sum().backward()
        xc = x.detach().cpu().requires_grad_(True)
        w, b = m.linear.weight.detach().cpu().requires_grad_(True), m.linear.bias.detach().cpu().requires_grad_(True)
        yr = ref_act(xc @ w.t() + b + xc)
        (yr ** 2).sum().backward()
        np.testing.assert_allclose(y.detach().cpu().numpy(), yr.detach().numpy(), rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(x.grad.cpu().numpy(), xc.grad.numpy(), rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(m.linear.weight.grad.cpu().numpy(), w.grad.numpy(), rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(m.linear.bias.grad.cpu().numpy(), b.grad.numpy(), rtol=1e-4, atol=1e-5)


# ---- --vanilla (reference train_mnist.py:351-357): plain PyTorch generator, outside the fused path ------------------
@pytest.mark.parametrize("script,argv", [
    ("train_mnist", ["--vanilla", "--synthetic", "128", "--synthetic_size", "12", "--num_epochs", "2", "--minibatch_size", "32",
                     "--p_hidden_dim", "32", "--q_hidden_dim", "32", "--seed", "0", "--learning_rate", "0.002"]),
    ("train_particles", ["--vanilla", "--synthetic", "64", "--synthetic-size", "12", "--num-epochs", "2", "--minibatch-size",
                         "32", "--p-hidden-dim", "32", "--q-hidden-dim", "32", "--fit-noise", "--seed", "0", "--learning-rate", "0.002"]),
    ("train_galaxy", ["--vanilla", "--synthetic", "48", "--synthetic_size", "8", "--num_epochs", "2", "--minibatch_size", "16",
                      "--p_hidden_dim", "32", "--q_hidden_dim", "32", "-z", "4", "--seed", "0", "--learning_rate", "0.002"]),
])
def test_vanilla_command_lines_run_the_plain_pytorch_generator(script, argv):
    _cuda()
    mod = _script(script)
    with contextlib.redirect_stdout(io.StringIO()) as buf:
        mod.main(argv)
    rows = [l.split("\t") for l in buf.getvalue().splitlines() if "\t" in l][1:]
    assert len(rows) == 4 and all(math.isfinite(float(r[-3])) for r in rows), rows
    assert float(rows[2][-3]) > float(rows[0][-3]), rows    # the training ELBO improves


# ---- Trainer: graph replay equals the eager step with in-kernel eps -----------------------------------------------------
def test_graphed_step_with_in_kernel_eps_matches_eager():
    import spatial_vae.models as M
    import torch.nn as nn
    from spatial_vae.trainer import Trainer
    dev = _cuda()
    SF = _sf()
    grid = O.make_grid(12, 12).to(dev)
    g = torch.Generator().manual_seed(4)
    ys = [((torch.rand(32, 144, generator=g) > 0.8).float() * torch.rand(32, 144, generator=g)).to(dev) for _ in range(4)]
    res = {}
    for mode in ("eager", "graph"):
        torch.manual_seed(5)
        with contextlib.redirect_stdout(io.StringIO()):
            p = M.SpatialGenerator(3, 64, num_layers=2, activation=nn.Tanh).to(dev)
            q = M.InferenceNetwork(144, 6, 64, num_layers=2, activation=nn.Tanh).to(dev)
        tr = Trainer(p, q, SF.StepSpec(family="mnist", theta_prior=0.8, precision="parity"), lr=1e-3, seed=77)
        outs = [(tr.step if mode == "eager" else tr.step_graphed)(grid, y).clone() for y in ys]
        torch.cuda.synchronize()
        res[mode] = (torch.stack(outs).cpu(), tr.flat.data.clone().cpu())
    np.testing.assert_allclose(res["eager"][0].numpy(), res["graph"][0].numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(res["eager"][1].numpy(), res["graph"][1].numpy(), rtol=1e-4, atol=1e-6)
