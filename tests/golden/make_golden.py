#!/usr/bin/env python
"""Generate tests/golden/*.npz by RUNNING THE REFERENCE (read-only at /root/reference).

Run once in the build container:  python tests/golden/make_golden.py
The reference ships no golden vectors of its own, so these files are
the parity pins for oracle/svae_oracle.py and, through it, for the CUDA path.  The
reference cannot travel to the GPU box; the .npz files do.

Each case stores: the module state_dicts ("p.<key>", "q.<key>"), the inputs (y, eps, ctf,
mask, ...), and what the reference returned (elbo, logp, kl, y_hat) plus the gradient of
-elbo w.r.t. every parameter ("gp.<key>", "gq.<key>").  eps is injected by patching
Tensor.normal_ for the one (B, I) draw eval_minibatch makes (train_mnist.py:38).
"""
import contextlib
import io
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

REF = os.environ.get("SVAE_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))

sys.path.insert(0, REF)
for name in ("skimage", "skimage.transform", "matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["skimage.transform"].resize = None
with contextlib.redirect_stdout(io.StringIO()):
    import train_mnist, train_particles, train_galaxy          # noqa: E402
    import spatial_vae.models as ref_models                      # noqa: E402
    import spatial_vae.ctf as ref_ctf                            # noqa: E402


@contextlib.contextmanager
def inject_eps(eps):
    orig = torch.Tensor.normal_

    def patched(self, *a, **k):
        if tuple(self.shape) == tuple(eps.shape):
            return self.copy_(eps)
        return orig(self, *a, **k)

    torch.Tensor.normal_ = patched
    try:
        yield
    finally:
        torch.Tensor.normal_ = orig


def build(P_in, z_dim, inf_dim, H, L, Hq, Lq, C, act=nn.Tanh, seed=0):
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        p = ref_models.SpatialGenerator(z_dim, H, n_out=C, num_layers=L, activation=act)
        q = ref_models.InferenceNetwork(P_in, inf_dim, Hq, num_layers=Lq, activation=act)
    return p, q


def grid_of(n, m):
    # verbatim recipe of train_mnist.py:316-320
    xgrid = np.linspace(-1, 1, m)
    ygrid = np.linspace(1, -1, n)
    x0, x1 = np.meshgrid(xgrid, ygrid)
    return torch.from_numpy(np.stack([x0.ravel(), x1.ravel()], 1)).float()


def pack(p, q, extra):
    d = {}
    for k, v in p.state_dict().items():
        d["p." + k] = v.detach().numpy().copy()
    for k, v in q.state_dict().items():
        d["q." + k] = v.detach().numpy().copy()
    for k, v in p.named_parameters():
        d["gp." + k] = (v.grad if v.grad is not None else torch.zeros_like(v)).numpy().copy()
    for k, v in q.named_parameters():
        d["gq." + k] = (v.grad if v.grad is not None else torch.zeros_like(v)).numpy().copy()
    for k, v in extra.items():
        d[k] = v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)
    return d


def mnist_case(name, rotate, translate, act=nn.Tanh, L=2, n=6, m=5, B=5, Z=3, seed=1):
    P = n * m
    I = Z + (1 if rotate else 0) + (2 if translate else 0)
    p, q = build(P, Z, I, 16, L, 12, 2, 1, act, seed)
    g = torch.Generator().manual_seed(100 + seed)
    y = (torch.rand(B, P, generator=g) > 0.7).float() * torch.rand(B, P, generator=g)
    eps = torch.randn(B, I, generator=g)
    x = grid_of(n, m)
    with inject_eps(eps):
        elbo, logp, kl, y_hat = train_mnist.eval_minibatch(x, y, p, q, rotate=rotate, translate=translate,
                                                          dx_scale=0.1, theta_prior=np.pi / 4)
    (-elbo).backward()
    d = pack(p, q, dict(y=y, eps=eps, grid=x, elbo=elbo, logp=logp, kl=kl, y_hat=y_hat,
                        n=n, m=m, rotate=int(rotate), translate=int(translate), L=L,
                        theta_prior=np.pi / 4, dx_scale=0.1,
                        act="tanh" if act is nn.Tanh else "leakyrelu"))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)


def particles_case(name, fit_noise=False, use_ctf=False, use_mask=False, augment=False, z_scale=1.0,
                   n=6, B=4, Z=2, seed=2, rotate=True, translate=True):
    P = n * n
    I = Z + int(rotate) + 2 * int(translate)
    C = 2 if fit_noise else 1
    p, q = build(P, Z, I, 16, 2, 12, 2, C, nn.Tanh, seed)
    g = torch.Generator().manual_seed(200 + seed)
    y = torch.randn(B, P, generator=g)
    eps = torch.randn(B, I, generator=g)
    x = grid_of(n, n)
    ctf = None
    if use_ctf:
        ctf = 0.05 * torch.randn(B, 1, n - 1, n - 1, generator=g)
    mask = None
    if use_mask:  # recipe of train_particles.py:387-396
        radius = n / 2
        yy, xx = np.ogrid[:n, :n]
        dist = np.sqrt((n / 2 - yy) ** 2 + (n / 2 - xx) ** 2)
        mask = (torch.from_numpy(dist) < radius).view(-1)
    extra = {}
    if augment:
        from PIL import Image
        np.random.seed(7)
        offset = np.random.uniform(0, 2 * np.pi, size=B)
        y_rot = y.clone()
        for i in range(B):
            im = Image.fromarray(y[i].view(n, n).numpy())
            im = im.rotate(360 * offset[i] / 2 / np.pi, resample=Image.BICUBIC)
            y_rot[i] = torch.from_numpy(np.array(im)).view(-1)
        extra = dict(theta_offset=offset.astype(np.float32), y_enc=y_rot)
        np.random.seed(7)  # the reference draws the same offsets
    with inject_eps(eps):
        elbo, logp, kl = train_particles.eval_minibatch(x, y, mask, ctf, p, q, rotate=rotate, translate=translate,
                                                       dx_scale=0.1, theta_prior=np.pi,
                                                       augment_rotation=augment, z_scale=z_scale)
    (-elbo).backward()
    extra.update(y=y, eps=eps, grid=x, elbo=elbo, logp=logp, kl=kl, n=n, z_scale=z_scale,
                 fit_noise=int(fit_noise), theta_prior=np.pi, dx_scale=0.1, rotate=int(rotate),
                 translate=int(translate))
    if ctf is not None:
        extra["ctf"] = ctf
    if mask is not None:
        extra["mask"] = mask
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **pack(p, q, extra))


def particles_option_case(name, seed, **opts):
    """--resid / --expand-coords / --bilinear / --softplus decoders and --resid encoders (train_particles.py:289-293,
    437-444): options the B200 kernels do not implement yet; the fixtures pin the oracle for when they do."""
    n, B, Z = 6, 4, 2
    P, I = n * n, Z + 3
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        p = ref_models.SpatialGenerator(Z, 16, n_out=1, num_layers=3, activation=nn.Tanh,
                                        softplus=opts.get("softplus", False), resid=opts.get("resid", False),
                                        expand_coords=opts.get("expand_coords", False),
                                        bilinear=opts.get("bilinear", False))
        q = ref_models.InferenceNetwork(P, I, 12, num_layers=3, activation=nn.Tanh, resid=opts.get("resid", False))
    g = torch.Generator().manual_seed(600 + seed)
    y = torch.randn(B, P, generator=g)
    eps = torch.randn(B, I, generator=g)
    x = grid_of(n, n)
    with inject_eps(eps):
        elbo, logp, kl = train_particles.eval_minibatch(x, y, None, None, p, q, rotate=True, translate=True,
                                                       dx_scale=0.1, theta_prior=np.pi)
    (-elbo).backward()
    extra = dict(y=y, eps=eps, grid=x, elbo=elbo, logp=logp, kl=kl, n=n, z_scale=1.0, theta_prior=np.pi, dx_scale=0.1,
                 **{"opt_" + k: int(v) for k, v in opts.items()})
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **pack(p, q, extra))


def galaxy_case(name, n=4, B=3, Z=4, L=3, seed=3, z_scale=1.0):
    P = n * n
    I = Z + 3
    p, q = build(P * 3, Z, I, 16, L, 20, 2, 3, nn.Tanh, seed)
    g = torch.Generator().manual_seed(300 + seed)
    y = torch.rand(B, P, 3, generator=g)
    eps = torch.randn(B, I, generator=g)
    x = grid_of(n, n)
    with inject_eps(eps):
        elbo, logp, kl, y_hat = train_galaxy.eval_minibatch(x, y, p, q, rotate=True, translate=True,
                                                           dx_scale=0.1, theta_prior=np.pi, z_scale=z_scale)
    (-elbo).backward()
    # display / generation helpers (train_galaxy.py:131-183; note the argument order q_net, p_net), with the
    # normal draws injected: r = eps for the posterior sample, z_rand for the prior sample
    z_rand = torch.randn(B, Z, generator=g)
    with torch.no_grad():
        with inject_eps(eps):
            display = train_galaxy.minibatch_for_display(x, y, q, p, rotate=True, translate=True, z_scale=0.8)
        with inject_eps(z_rand):
            generated = train_galaxy.random_minibatch_generator(x, y, p, Z, z_scale=0.8)
    d = pack(p, q, dict(y=y, eps=eps, grid=x, elbo=elbo, logp=logp, kl=kl, y_hat=y_hat, n=n, L=L,
                        z_scale=z_scale, theta_prior=np.pi, dx_scale=0.1, z_rand=z_rand, display=display,
                        generated=generated))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)


def trajectory_case(name, steps=10, n=6, m=6, B=8, Z=3, seed=4):
    """10 reference train steps (eval_minibatch, backward, Adam.step, zero_grad) with the
    reference's optimiser construction (train_mnist.py:387-392, 147-150)."""
    P = n * m
    I = Z + 3
    p, q = build(P, Z, I, 16, 2, 12, 2, 1, nn.Tanh, seed)
    init = {"p." + k: v.detach().numpy().copy() for k, v in p.state_dict().items()}
    init.update({"q." + k: v.detach().numpy().copy() for k, v in q.state_dict().items()})
    optim = torch.optim.Adam(list(p.parameters()) + list(q.parameters()), lr=1e-4)
    g = torch.Generator().manual_seed(400 + seed)
    x = grid_of(n, m)
    ys, es, elbos = [], [], []
    for _ in range(steps):
        y = (torch.rand(B, P, generator=g) > 0.7).float() * torch.rand(B, P, generator=g)
        eps = torch.randn(B, I, generator=g)
        with inject_eps(eps):
            elbo, _, _, _ = train_mnist.eval_minibatch(x, y, p, q, rotate=True, translate=True,
                                                       dx_scale=0.1, theta_prior=np.pi / 4)
        (-elbo).backward()
        optim.step()
        optim.zero_grad()
        ys.append(y.numpy())
        es.append(eps.numpy())
        elbos.append(float(elbo))
    d = {"init." + k: v for k, v in init.items()}
    for k, v in p.state_dict().items():
        d["final.p." + k] = v.detach().numpy().copy()
    for k, v in q.state_dict().items():
        d["final.q." + k] = v.detach().numpy().copy()
    d.update(ys=np.stack(ys), eps=np.stack(es), elbos=np.array(elbos), grid=x.numpy(), n=n, m=m,
             theta_prior=np.pi / 4, dx_scale=0.1, lr=1e-4)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)


def decoder_case(name, n=5, m=7, B=3, Z=4, L=3, C=2, seed=5):
    """Module-level SpatialGenerator.forward on explicit coordinates (models.py:90-132)."""
    p, _ = build(n * m, Z, Z, 16, L, 8, 1, C, nn.Tanh, seed)
    g = torch.Generator().manual_seed(500 + seed)
    x = torch.randn(B, n * m, 2, generator=g)
    z = torch.randn(B, Z, generator=g)
    y = p(x, z)
    d = {"p." + k: v.detach().numpy().copy() for k, v in p.state_dict().items()}
    d.update(x=x.numpy(), z=z.numpy(), y_hat=y.detach().numpy(), L=L)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)


def ctf_case(name, N=3, n=7, m=7):
    import pandas as pd
    rng = np.random.default_rng(11)
    tab = pd.DataFrame(dict(defocus=rng.uniform(1, 3, N), cs=np.full(N, 2.7), voltage=np.full(N, 300.0),
                            apix=np.full(N, 2.5), bfactor=np.full(N, 100.0), ampcont=np.full(N, 10.0),
                            dfdiff=np.zeros(N), dfang=rng.uniform(0, 180, N)))
    k = ref_ctf.ctf_filter(tab, n, m)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), kernels=k, n=n, m=m,
                        **{c: tab[c].to_numpy() for c in tab.columns})


def cli_defaults_case(name):
    """Every command-line flag of the three reference drivers with its default value (repr), read from the reference's
    own argparse parsers: train_mnist.mnist_arguments, train_galaxy.galaxy_arguments, and the parser that
    train_particles.main builds inline (train_particles.py:277-318; its block is executed as written)."""
    import argparse
    import json
    out = {}
    old = sys.argv
    try:
        sys.argv = ["train_mnist.py"]
        out["train_mnist"] = {k: repr(v) for k, v in vars(train_mnist.mnist_arguments()).items()}
        sys.argv = ["train_galaxy.py", "train.npy", "test.npy"]
        out["train_galaxy"] = {k: repr(v) for k, v in vars(train_galaxy.galaxy_arguments()).items()}
    finally:
        sys.argv = old
    src = open(os.path.join(REF, "train_particles.py")).read()
    block = src[src.index("parser = argparse.ArgumentParser"):src.index("args = parser.parse_args()")]
    block = "\n".join(l[4:] if l.startswith("    ") else l for l in block.splitlines())
    ns = {"argparse": argparse, "np": np}
    exec(block, ns)
    out["train_particles"] = {k: repr(v) for k, v in vars(ns["parser"].parse_args(["train.npy", "test.npy"])).items()}
    with open(os.path.join(OUT, name + ".json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


def driver_signatures_case(name):
    """inspect.signature of every function the three reference driver scripts define themselves."""
    import inspect
    import json
    out = {}
    for mod in (train_mnist, train_particles, train_galaxy):
        for fn, obj in inspect.getmembers(mod, inspect.isfunction):
            if obj.__module__ == mod.__name__:
                out[f"{mod.__name__}.{fn}"] = str(inspect.signature(obj))
    with open(os.path.join(OUT, name + ".json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


def ingest_case(name):
    """The reference's host-side ingest helpers on fixed inputs: image.downsample / crop / normalize,
    ctf.compute_2d_ctf with astigmatism, and MRC files as mrc.write produces them (default header; make_header +
    extended header, int16 voxels)."""
    import io as _io
    import spatial_vae.image as ref_image
    import spatial_vae.mrc as ref_mrc
    rng = np.random.default_rng(17)
    stack = rng.standard_normal((3, 12, 10)).astype(np.float32)
    out = {"stack": stack,
           "down_factor2": ref_image.downsample(stack, factor=2),
           "down_shape_5x7": ref_image.downsample(stack, shape=(5, 7)),
           "down_2d": ref_image.downsample(stack[0], factor=1.5),
           "crop8": ref_image.crop(stack, 8),
           "norm_default": ref_image.normalize(stack),
           "norm_r3": ref_image.normalize(stack, radius=3)}
    freqs = np.stack([g.ravel() for g in np.meshgrid(np.fft.fftfreq(9), np.fft.fftfreq(11), indexing="ij")], 1) / 2.5
    out["freqs"] = freqs
    out["ctf_astig"] = ref_ctf.compute_2d_ctf(freqs, 21000.0, 18000.0, 0.6, 300.0, 2.7, 0.1, 120.0)
    out["ctf_no_b"] = ref_ctf.compute_2d_ctf(freqs, 15000.0, 15000.0, 0.0, 200.0, 2.0, 0.07)
    buf = _io.BytesIO()
    ref_mrc.write(buf, stack)
    out["mrc_default"] = np.frombuffer(buf.getvalue(), dtype=np.uint8)
    vol = rng.integers(-500, 500, size=(2, 4, 6)).astype(np.int16)
    hdr = ref_mrc.make_header(vol.shape, (6.0, 4.0, 2.0), (90.0, 90.0, 90.0), mz=2, dtype=np.int16, exthd_size=16)
    buf = _io.BytesIO()
    ref_mrc.write(buf, vol, header=hdr, extended_header=bytes(range(16)))
    out["vol_int16"] = vol
    out["mrc_int16_ext"] = np.frombuffer(buf.getvalue(), dtype=np.uint8)
    buf = _io.BytesIO()
    ref_mrc.write(buf, stack[:1], ax=25.0, ay=30.0, az=1.0, alpha=90.0, beta=90.0, gamma=90.0)
    out["mrc_single_cell"] = np.frombuffer(buf.getvalue(), dtype=np.uint8)
    out["modes"] = np.array([ref_mrc.get_mode(np.dtype(t)) for t in ("int8", "int16", "float32", "complex64", "uint16")])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)


def pickle_case(name, seed, **options):
    """Whole-module pickles exactly as the reference writes them (torch.save(p_net, path), misc_tools.py:93-99,
    train_particles.py:530-543) plus what the reference modules return on fixed inputs: the interop fixture for
    loading reference-trained .sav files into the B200 package."""
    torch.manual_seed(seed)
    n, Z, H, Hq = 6, 3, 16, 12
    with contextlib.redirect_stdout(io.StringIO()):
        p = ref_models.SpatialGenerator(Z, H, n_out=2, num_layers=3, activation=nn.Tanh, **options)
        q = ref_models.InferenceNetwork(n * n, Z + 3, Hq, num_layers=3, activation=nn.Tanh,
                                        resid=options.get("resid", False))
    p.eval().cpu()
    q.eval().cpu()
    torch.save(p, os.path.join(OUT, name + "_generator.sav"))
    torch.save(q, os.path.join(OUT, name + "_inference.sav"))
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.rand(4, n * n, 2, generator=g) * 2 - 1
    z = torch.randn(4, Z, generator=g)
    y = torch.randn(4, n * n, generator=g)
    with torch.no_grad():
        y_hat = p(x, z)
        z_mu, z_logstd = q(y)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), x=x.numpy(), z=z.numpy(), y=y.numpy(), y_hat=y_hat.numpy(),
                        z_mu=z_mu.numpy(), z_logstd=z_logstd.numpy())


if __name__ == "__main__":
    cli_defaults_case("cli_defaults")
    driver_signatures_case("driver_signatures")
    ingest_case("ingest")
    pickle_case("ref_pickle_plain", 41)
    pickle_case("ref_pickle_options", 42, softplus=True, resid=True, expand_coords=True, bilinear=True)
    mnist_case("mnist_rt", True, True)
    mnist_case("mnist_r", True, False, seed=11)
    mnist_case("mnist_t", False, True, seed=12)
    mnist_case("mnist_none", False, False, seed=13)
    mnist_case("mnist_leaky_L3", True, True, act=nn.LeakyReLU, L=3, seed=14)
    particles_case("particles_plain")
    particles_case("particles_fitnoise", fit_noise=True, seed=21)
    particles_case("particles_ctf", use_ctf=True, seed=22)
    particles_case("particles_mask", use_mask=True, seed=23)
    particles_case("particles_augment", augment=True, seed=24)
    particles_case("particles_zscale0", z_scale=0.0, seed=25)
    particles_case("particles_t_only", rotate=False, fit_noise=True, z_scale=0.6, seed=26)
    particles_case("particles_r_only", translate=False, use_mask=True, seed=27)
    particles_option_case("particles_opt_resid", 31, resid=True)
    particles_option_case("particles_opt_expand", 32, expand_coords=True)
    particles_option_case("particles_opt_bilinear", 33, bilinear=True)
    particles_option_case("particles_opt_softplus", 34, softplus=True)
    particles_option_case("particles_opt_all", 35, resid=True, expand_coords=True, bilinear=True, softplus=True)
    galaxy_case("galaxy_rgb")
    trajectory_case("mnist_adam10")
    decoder_case("decoder_module")
    ctf_case("ctf_kernels")
    print("wrote", sorted(f for f in os.listdir(OUT) if f.endswith(".npz")))
