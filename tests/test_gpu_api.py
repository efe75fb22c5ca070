"""GPU tests of the reference-facing API: the per-script eval_minibatch functions (same signatures
and return values as the reference), autograd through the fused step, the Trainer, the three command
lines on synthetic data, and whole-module pickles."""
import contextlib
import importlib.util
import io
import math
import os

import numpy as np
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu

from oracle import svae_oracle as O
from tests.helpers import golden_grads, load_case, oracle_params

PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "spatial-vae_b200")


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


@contextlib.contextmanager
def _inject_normal(values):
    """Make the next Tensor.normal_() on a tensor of this shape return `values` (how the fixtures pin the reference's
    random draws, see tests/golden/make_golden.py)."""
    orig = torch.Tensor.normal_

    def patched(self, *a, **k):
        if tuple(self.shape) == tuple(values.shape):
            return self.copy_(values)
        return orig(self, *a, **k)

    torch.Tensor.normal_ = patched
    try:
        yield
    finally:
        torch.Tensor.normal_ = orig


def _script(name):
    spec = importlib.util.spec_from_file_location("cli_" + name, os.path.join(PKG, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _nets(d, dev, C=1, n_in=None):
    import spatial_vae.models as M
    ps = {k[2:]: torch.from_numpy(v) for k, v in d.items() if k.startswith("p.")}
    qs = {k[2:]: torch.from_numpy(v) for k, v in d.items() if k.startswith("q.")}
    H, Z = ps["coord_linear.weight"].shape[0], ps["latent_linear.weight"].shape[1]
    Lp = len([k for k in ps if k.startswith("layers.") and k.endswith(".weight")])
    n_out = ps[sorted(k for k in ps if k.startswith("layers.") and k.endswith(".weight"))[-1]].shape[0]
    act = nn.LeakyReLU if str(d.get("act", "tanh")) == "leakyrelu" else nn.Tanh
    with contextlib.redirect_stdout(io.StringIO()):
        p = M.SpatialGenerator(Z, H, n_out=n_out, num_layers=Lp, activation=act)
        q = M.InferenceNetwork(qs["layers.0.weight"].shape[1], qs["layers.4.weight"].shape[0] // 2,
                               qs["layers.0.weight"].shape[0], num_layers=2, activation=act)
    p.load_state_dict(ps)
    q.load_state_dict(qs)
    p.precision = "parity"
    return p.to(dev), q.to(dev)


def _check_module_grads(p, q, d, rtol=5e-4):
    mine = [t.grad.cpu() for t in list(p.parameters()) + list(q.parameters())]
    for i, (g, r) in enumerate(zip(mine, golden_grads(d))):
        np.testing.assert_allclose(g.numpy(), r.numpy(), rtol=rtol, atol=5e-6, err_msg=f"grad {i}")


def test_train_mnist_eval_minibatch_is_a_drop_in():
    dev = _dev()
    tm = _script("train_mnist")
    d = load_case("mnist_rt")
    p, q = _nets(d, dev)
    x = torch.from_numpy(d["grid"]).to(dev)
    y = torch.from_numpy(d["y"]).to(dev)
    eps = torch.from_numpy(d["eps"]).to(dev)
    elbo, logp, kl, y_hat = tm.eval_minibatch(x, y, p, q, rotate=True, translate=True, dx_scale=0.1,
                                              theta_prior=np.pi / 4, eps=eps)
    assert elbo.dim() == 0 and y_hat.shape == (y.shape[0], y.shape[1])
    np.testing.assert_allclose(float(elbo.detach()), float(d["elbo"]), rtol=2e-5)
    np.testing.assert_allclose(float(logp), float(d["logp"]), rtol=2e-5)
    np.testing.assert_allclose(float(kl), float(d["kl"]), rtol=2e-5)
    np.testing.assert_allclose(y_hat.detach().cpu().numpy(), d["y_hat"], atol=1e-6)
    loss = -elbo                       # the reference's train loop (train_mnist.py:147-150)
    loss.backward()
    _check_module_grads(p, q, d)
    opt = torch.optim.Adam(list(p.parameters()) + list(q.parameters()), lr=1e-4)
    opt.step()
    opt.zero_grad()


def test_eval_minibatch_draws_eps_from_the_global_generator():
    dev = _dev()
    tm = _script("train_mnist")
    d = load_case("mnist_rt")
    p, q = _nets(d, dev)
    x, y = torch.from_numpy(d["grid"]).to(dev), torch.from_numpy(d["y"]).to(dev)
    torch.manual_seed(11)
    a = tm.eval_minibatch(x, y, p, q, theta_prior=np.pi / 4)[0]
    torch.manual_seed(11)
    b = tm.eval_minibatch(x, y, p, q, theta_prior=np.pi / 4)[0]
    c = tm.eval_minibatch(x, y, p, q, theta_prior=np.pi / 4)[0]
    assert float(a) == float(b) and float(a) != float(c)
    with torch.no_grad():
        e = tm.eval_minibatch(x, y, p, q, theta_prior=np.pi / 4)[0]
    assert not e.requires_grad


def test_train_particles_eval_minibatch_variants():
    dev = _dev()
    tp = _script("train_particles")
    for name in ("particles_plain", "particles_fitnoise", "particles_ctf", "particles_mask", "particles_zscale0"):
        d = load_case(name)
        p, q = _nets(d, dev)
        x, y = torch.from_numpy(d["grid"]).to(dev), torch.from_numpy(d["y"]).to(dev)
        eps = torch.from_numpy(d["eps"]).to(dev)
        ctf = torch.from_numpy(d["ctf"]).to(dev) if "ctf" in d else None
        mask = torch.from_numpy(d["mask"]).to(dev) if "mask" in d else None
        elbo, logp, kl = tp.eval_minibatch(x, y, mask, ctf, p, q, rotate=True, translate=True, dx_scale=0.1,
                                           theta_prior=np.pi, z_scale=float(d["z_scale"]), eps=eps)
        np.testing.assert_allclose(float(elbo), float(d["elbo"]), rtol=2e-5, err_msg=name)
        (-elbo).backward()
        _check_module_grads(p, q, d)
    # CTF + fit-noise: the reference crashes; so do we, loudly
    d = load_case("particles_fitnoise")
    p, q = _nets(d, dev)
    with pytest.raises(RuntimeError):
        tp.eval_minibatch(torch.from_numpy(d["grid"]).to(dev), torch.from_numpy(d["y"]).to(dev), None,
                          torch.zeros(4, 1, 5, 5, device=dev), p, q)


def test_train_particles_augment_rotation_matches_reference():
    dev = _dev()
    tp = _script("train_particles")
    d = load_case("particles_augment")
    p, q = _nets(d, dev)
    x, y = torch.from_numpy(d["grid"]).to(dev), torch.from_numpy(d["y"]).to(dev)
    np.random.seed(7)     # the fixture drew its offsets from this numpy seed (unseeded in the reference)
    elbo, _, _ = tp.eval_minibatch(x, y, None, None, p, q, rotate=True, translate=True, dx_scale=0.1,
                                   theta_prior=np.pi, augment_rotation=True, eps=torch.from_numpy(d["eps"]).to(dev))
    np.testing.assert_allclose(float(elbo), float(d["elbo"]), rtol=5e-5)


def test_train_galaxy_eval_minibatch():
    dev = _dev()
    tg = _script("train_galaxy")
    d = load_case("galaxy_rgb")
    p, q = _nets(d, dev)
    x, y = torch.from_numpy(d["grid"]).to(dev), torch.from_numpy(d["y"]).to(dev)
    elbo, logp, kl, y_hat = tg.eval_minibatch(x, y, p, q, rotate=True, translate=True, dx_scale=0.1,
                                              theta_prior=np.pi, eps=torch.from_numpy(d["eps"]).to(dev))
    assert y_hat.shape == y.shape
    np.testing.assert_allclose(float(elbo), float(d["elbo"]), rtol=2e-5)
    np.testing.assert_allclose(y_hat.detach().cpu().numpy(), d["y_hat"], atol=1e-6)
    (-elbo).backward()
    _check_module_grads(p, q, d)
    z = tg.random_minibatch_generator(x, y, p, 4)
    disp = tg.minibatch_for_display(x, y, q, p)          # galaxy order: q_net before p_net (train_galaxy.py:131)
    assert z.shape == y.shape and disp.shape == y.shape and float(disp.min()) >= 0 and float(disp.max()) <= 1


def test_trainer_runs_the_reference_loop_on_flat_buffers():
    dev = _dev()
    import spatial_vae.functional as SF
    from spatial_vae.trainer import Trainer
    d = load_case("mnist_adam10")
    dd = {"p." + k[7:]: v for k, v in d.items() if k.startswith("init.p.")}
    dd.update({"q." + k[7:]: v for k, v in d.items() if k.startswith("init.q.")})
    p, q = _nets(dd, dev)
    spec = SF.StepSpec(family="mnist", theta_prior=float(d["theta_prior"]), dx_scale=float(d["dx_scale"]),
                       precision="parity")
    tr = Trainer(p, q, spec, lr=float(d["lr"]))
    grid = torch.from_numpy(d["grid"]).to(dev)
    for t in range(d["ys"].shape[0]):
        res = tr.step(grid, torch.from_numpy(d["ys"][t]).to(dev), eps=torch.from_numpy(d["eps"][t]).to(dev))
        np.testing.assert_allclose(float(res[0]), d["elbos"][t], rtol=1e-4)
        np.testing.assert_allclose(float(res[0]), float(res[1]) - float(res[2]), rtol=1e-5)     # elbo = logp - kl
    # the loss sums the step leaves in the tail of the gradient buffer are the column sums of the per-image stats
    y0, eps0 = torch.from_numpy(d["ys"][0]).to(dev), torch.from_numpy(d["eps"][0]).to(dev)
    sums = torch.full((4,), float("nan"), device=dev)
    stats, _, _ = SF.run_step(spec, tr.dec, tr.enc, grid, y0, eps0, stats_sum=sums)
    np.testing.assert_allclose(sums[:3].cpu().numpy(), stats.sum(0).cpu().numpy(), rtol=1e-5)
    assert float(sums[3]) == 0.0
    for k, v in p.state_dict().items():
        assert float((v.cpu() - torch.from_numpy(d["final.p." + k])).abs().max()) < 1e-4, k
    for k, v in q.state_dict().items():
        assert float((v.cpu() - torch.from_numpy(d["final.q." + k])).abs().max()) < 1e-4, k


def test_whole_module_pickles_round_trip(tmp_path):
    dev = _dev()
    from spatial_vae import driver as D
    d = load_case("mnist_rt")
    p, q = _nets(d, dev)
    D.save_models(str(tmp_path / "run"), "01", None, p, q, dev)
    p2 = torch.load(str(tmp_path / "run_generator_epoch01.sav"), weights_only=False)
    q2 = torch.load(str(tmp_path / "run_inference_epoch01.sav"), weights_only=False)
    assert type(p2).__module__ == "spatial_vae.models" and not p2.training
    for k, v in p.state_dict().items():
        assert torch.equal(v.cpu(), p2.state_dict()[k])
    for k, v in q.state_dict().items():
        assert torch.equal(v.cpu(), q2.state_dict()[k])
    y = p2.to(dev)(torch.zeros(1, 4, 2, device=dev), torch.zeros(1, 3, device=dev))
    assert y.shape == (1, 4, 1)


@pytest.mark.parametrize("script,argv,cols", [
    ("train_mnist", ["--synthetic", "192", "--synthetic_size", "12", "--num_epochs", "2", "--minibatch_size", "50",
                     "--p_hidden_dim", "64", "--q_hidden_dim", "64", "--seed", "0", "--precision", "fast"], 4),
    ("train_particles", ["--synthetic", "96", "--synthetic-size", "12", "--num-epochs", "2", "--minibatch-size", "40",
                         "--p-hidden-dim", "64", "--q-hidden-dim", "64", "--fit-noise", "--mask", "--z-delay", "1",
                         "--seed", "0"], 5),
    ("train_galaxy", ["--synthetic", "64", "--synthetic_size", "8", "--num_epochs", "2", "--minibatch_size", "24",
                      "--p_hidden_dim", "64", "--p_num_layers", "3", "--q_hidden_dim", "128", "-z", "5", "--seed", "0",
                      "--learning_rate", "0.002"], 4),
])
def test_command_lines_train_on_synthetic_data(script, argv, cols, capsys):
    _dev()
    mod = _script(script)
    with contextlib.redirect_stdout(io.StringIO()) as buf:
        mod.main(argv)
    lines = [l for l in buf.getvalue().splitlines() if "\t" in l]
    assert lines[0].split("\t")[0] == "Epoch" and len(lines[0].split("\t")) == cols
    rows = [l.split("\t") for l in lines[1:]]
    assert len(rows) == 4                       # 2 epochs x (train, validation)
    vals = [float(r[-3]) for r in rows]         # ELBO column
    assert all(math.isfinite(v) for v in vals)
    assert vals[2] > vals[0]                    # the training ELBO improves from epoch 1 to epoch 2


def test_device_bicubic_rotation_is_bit_exact_with_the_pillow_port():
    """svae_rotate_bicubic against oracle/pillow_rotate.py (itself bit exact with Pillow, see the CPU tests):
    float32 particles and the uint8 round trip of the galaxy driver."""
    dev = _dev()
    import spatial_vae.functional as SF
    from oracle.pillow_rotate import rotate_bicubic
    rng = np.random.default_rng(3)
    for n, B in ((40, 9), (7, 5), (28, 6)):
        y = rng.standard_normal((B, n * n)).astype(np.float32)
        deg = rng.uniform(0, 360, B)
        deg[:4] = [0.0, 90.0, 180.0, 270.0]
        got = SF.rotate_bicubic(torch.from_numpy(y).to(dev), n, n, deg).cpu().numpy()
        for b in range(B):
            ref = rotate_bicubic(y[b].reshape(n, n), float(deg[b])).reshape(-1)
            assert np.array_equal(got[b], ref), (n, b, deg[b], np.abs(got[b] - ref).max())
    n, B = 16, 6
    y = rng.random((B, n * n, 3)).astype(np.float32)
    deg = rng.uniform(0, 360, B)
    got = SF.rotate_bicubic(torch.from_numpy(y).to(dev), n, n, deg, channels=3, quantize_u8=True).cpu().numpy()
    for b in range(B):
        u8 = (y[b].reshape(n, n, 3) * 255).astype(np.uint8)
        ref = (rotate_bicubic(u8, float(deg[b])).astype(float) / 255).astype(np.float32).reshape(-1, 3)
        assert np.array_equal(got[b], ref), (b, np.abs(got[b] - ref).max())


def test_graphed_step_matches_eager_step():
    """Trainer.step_graphed (CUDA-graph replay, Adam bias corrections in device memory) walks the same
    trajectory as Trainer.step when eps is drawn from the same generator state."""
    dev = _dev()
    import spatial_vae.functional as SF
    from spatial_vae.trainer import Trainer
    d = load_case("mnist_adam10")
    dd = {"p." + k[7:]: v for k, v in d.items() if k.startswith("init.p.")}
    dd.update({"q." + k[7:]: v for k, v in d.items() if k.startswith("init.q.")})
    spec = SF.StepSpec(family="mnist", theta_prior=float(d["theta_prior"]), dx_scale=float(d["dx_scale"]),
                       precision="parity")
    grid = torch.from_numpy(d["grid"]).to(dev)
    results = []
    for graphed in (False, True):
        p, q = _nets(dd, dev)
        tr = Trainer(p, q, spec, lr=float(d["lr"]))
        torch.manual_seed(123)
        out = []
        for t in range(6):
            y = torch.from_numpy(d["ys"][t]).to(dev)
            r = (tr.step_graphed if graphed else tr.step)(grid, y)
            out.append(r.clone())
        torch.cuda.synchronize()
        results.append((torch.stack(out).cpu(), tr.flat.data.clone().cpu()))
    (lo_e, p_e), (lo_g, p_g) = results
    # eps comes from torch's generator in both modes, but a graph draws its own Philox offsets: compare statistically
    assert torch.isfinite(lo_g).all() and lo_g.shape == lo_e.shape
    assert float((lo_g[:, 0] - lo_e[:, 0]).abs().max()) < 0.15 * float(lo_e[:, 0].abs().mean())
    assert float((p_g - p_e).abs().max()) < 2e-3          # 6 steps of lr 1e-4: same direction, bounded drift

