"""CPU tests of the host-side pieces around the hot path: CLI flag surfaces, CTF kernels, MRC IO."""
import io
import math
import os
import sys

import numpy as np
import pytest

from tests.helpers import load_case

PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "spatial-vae_b200")


def test_ctf_kernels_match_reference_fixture():
    import pandas as pd
    import spatial_vae.ctf as C
    d = load_case("ctf_kernels")
    table = pd.DataFrame({k: d[k] for k in C.CTF_COLUMNS})
    k = C.ctf_filter(table, int(d["n"]), int(d["m"]))
    np.testing.assert_allclose(k, d["kernels"], rtol=1e-5, atol=1e-7)


def test_ctf_table_parser(tmp_path):
    import spatial_vae.ctf as C
    path = tmp_path / "ctf.txt"
    path.write_text("1.5 2.7 300 2.5 100 10 0 45\n2.0  2.7 300 2.5 100 10 0 90\n")
    t = C.parse_ctf(str(path))
    assert list(t.columns) == C.CTF_COLUMNS and len(t) == 2 and t.defocus[1] == 2.0


def test_mrc_roundtrip_and_crop():
    import spatial_vae.mrc as M
    from spatial_vae.image import crop
    a = np.random.default_rng(0).random((4, 6, 8)).astype(np.float32)
    buf = io.BytesIO()
    M.write(buf, a)
    b, header, ext = M.parse(buf.getvalue())
    assert np.array_equal(a, b) and (header.nx, header.ny, header.nz, header.mode) == (8, 6, 4, 2) and ext == b""
    assert crop(a, 4).shape == (4, 4, 4) and np.array_equal(crop(a, 4), a[:, 1:5, 2:6])


def _load_script(name):
    import importlib.util
    spec = importlib.util.spec_from_file_location("cli_" + name, os.path.join(PKG, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_cli_flags_keep_reference_names_and_defaults():
    mn = _load_script("train_mnist")
    a = mn.mnist_arguments([])
    assert (a.z_dim, a.p_hidden_dim, a.q_hidden_dim, a.num_layers, a.minibatch_size) == (2, 500, 500, 2, 100)
    assert a.theta_prior == pytest.approx(math.pi / 4) and a.dx_scale == 0.1 and a.learning_rate == 1e-4
    assert a.dataset == "mnist-rotated-translated" and a.device == -2 and a.val_split == 50
    b = mn.mnist_arguments(["--z-dim", "7", "--no_rotate", "--minibatch-size", "64"])   # both spellings
    assert b.z_dim == 7 and b.no_rotate and b.minibatch_size == 64

    pt = _load_script("train_particles")
    c = pt.parse(["tr.npy", "te.npy", "--fit-noise", "--augment-rotation", "--minibatch-size", "512"])
    assert c.fit_noise and c.augment_rotation and c.minibatch_size == 512 and c.theta_prior == pytest.approx(math.pi)
    assert (c.p_hidden_dim, c.p_num_layers, c.q_hidden_dim, c.q_num_layers, c.z_delay) == (500, 2, 500, 2, 0)

    gx = _load_script("train_galaxy")
    g = gx.galaxy_arguments(["tr.npy", "te.npy", "-z", "20", "--p_hidden_dim", "1000", "--p_num_layers", "4"])
    assert (g.z_dim, g.p_hidden_dim, g.p_num_layers, g.q_hidden_dim) == (20, 1000, 4, 5000)
    assert g.theta_prior == pytest.approx(math.pi) and g.activation == "tanh"
    # the scripts export the reference's function names
    for mod in (mn, pt, gx):
        assert callable(mod.eval_minibatch) and callable(mod.main)


def test_activation_flag_quirks():
    import torch.nn as nn
    from spatial_vae import driver as D
    assert D.activation_from_flag("relu", "mnist") is nn.LeakyReLU      # reference train_mnist.py:344-348
    assert D.activation_from_flag("relu", "galaxy") is nn.ReLU          # reference train_galaxy.py:431-432
    assert D.activation_from_flag("leakyrelu", "galaxy") is nn.Tanh     # typo at train_galaxy.py:429


def test_bench_flop_model_matches_survey_table():
    """bench.py's algorithmic FLOPs per image per train step against SURVEY.md section 8(d)
    (C1 1.187, C2 1.188, C3 2.426, C4 74.371, C5 2.430 GFLOP): the roofline fractions rest on these."""
    sys.path.insert(0, os.path.dirname(PKG))
    import bench
    expect = {"c1": 1.187e9, "c2": 1.188e9, "c3": 2.426e9, "c4": 74.371e9, "c5": 2.430e9}
    for name, gf in expect.items():
        got = bench.flops_per_image_train(bench.CONFIGS[name])
        assert abs(got - gf) / gf < 2e-3, (name, got, gf)


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` needs no GPU and prints the contract's line: the unmodified reference when
    baseline/_ref exists (oracle/make_ref.py copies it where /root/reference is present), else the oracle port."""
    import json
    import subprocess
    root = os.path.dirname(PKG)
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--config", "c1",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "images/s" and d["value"] > 0
    have_ref = os.path.isdir(os.path.join(root, "baseline", "_ref", "spatial_vae"))
    assert d["cpu_baseline"]["kind"] == ("reference" if have_ref else "port") and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["config"]["images_per_step"] == 100          # C1 runs at its real minibatch on the CPU


def test_particle_preprocessing_matches_reference_recipes(tmp_path):
    """normalize / mask / stack loading of the particle CLI against the reference's inline code
    (train_particles.py:248-255, 339-347, 387-396), restated here."""
    import torch
    import spatial_vae.mrc as M
    from spatial_vae import driver as D
    rng = np.random.default_rng(2)
    imgs = (rng.standard_normal((5, 8, 6)) * 3 + 10).astype(np.float32)
    n, m = imgs.shape[1:]
    mu = imgs.reshape(-1, n * m).mean(1)
    std = imgs.reshape(-1, n * m).std(1)
    ref = (imgs - mu[:, np.newaxis, np.newaxis]) / std[:, np.newaxis, np.newaxis]
    got = D.normalize_particles(imgs)
    assert got.dtype == ref.dtype and np.array_equal(got, ref)
    ints = (imgs * 10).astype(np.int16)                    # MRC mode 1 stacks are integers
    assert np.allclose(D.normalize_particles(ints).reshape(5, -1).std(1), 1.0)

    radius = min(n, m) / 2
    y_grid, x_grid = np.ogrid[:n, :m]
    center = np.array([n / 2, m / 2])
    dist = np.sqrt((center[0] - y_grid) ** 2 + (center[1] - x_grid) ** 2)
    assert torch.equal(D.circular_mask(n, m), (torch.from_numpy(dist) < radius).view(-1))

    np.save(tmp_path / "stack.npy", imgs)
    assert np.array_equal(D.load_particle_stack(str(tmp_path / "stack.npy")), imgs)
    with open(tmp_path / "stack.mrcs", "wb") as f:
        M.write(f, imgs)
    assert np.array_equal(D.load_particle_stack(str(tmp_path / "stack.mrcs")), imgs)


def test_command_line_flags_and_defaults_match_the_reference():
    """tests/golden/cli_defaults.json holds every flag of the reference's three argparse parsers with its default
    (written by tests/golden/make_golden.py from the reference itself): the mirrors accept the same flags with the same
    defaults, plus five of their own."""
    import importlib.util
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref = json.load(open(os.path.join(root, "tests", "golden", "cli_defaults.json")))
    parsers = {"train_mnist": ("mnist_arguments", []), "train_particles": ("parse", ["train.npy", "test.npy"]),
               "train_galaxy": ("galaxy_arguments", ["train.npy", "test.npy"])}
    for name, (fn, argv) in parsers.items():
        spec = importlib.util.spec_from_file_location("cli_defaults_" + name,
                                                      os.path.join(root, "spatial-vae_b200", name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mine = {k: repr(v) for k, v in vars(getattr(mod, fn)(argv)).items()}
        assert ref[name], name
        for k, v in ref[name].items():
            assert mine.get(k) == v, (name, k, v, mine.get(k))
        assert sorted(set(mine) - set(ref[name])) == ["precision", "seed", "synthetic", "synthetic_size", "yes"]


def test_pickles_written_here_load_in_the_reference(tmp_path):
    """The other direction of the .sav interop: whole-module pickles written by this package (driver.save_models)
    unpickle into the REFERENCE's classes and evaluate there (needs the reference checkout, so it only runs in the
    build container); outputs are compared with the oracle on the same parameters."""
    import contextlib
    import io
    import os
    import subprocess
    import sys
    import torch
    ref = os.environ.get("SVAE_REFERENCE", "/root/reference")
    if not os.path.isdir(os.path.join(ref, "spatial_vae")):
        pytest.skip("reference checkout not available")
    import spatial_vae.models as M
    from spatial_vae import driver as D
    from oracle import svae_oracle as O
    torch.manual_seed(9)
    with contextlib.redirect_stdout(io.StringIO()):
        p = M.SpatialGenerator(3, 16, n_out=2, num_layers=3, softplus=True, resid=True, expand_coords=True, bilinear=True)
        q = M.InferenceNetwork(36, 6, 12, num_layers=3, resid=True)
    p.precision = "parity"
    D.save_models(str(tmp_path / "run"), "07", None, p, q, torch.device("cpu"))
    x, z, y = torch.rand(2, 36, 2) * 2 - 1, torch.randn(2, 3), torch.randn(2, 36)
    torch.save({"x": x, "z": z, "y": y}, tmp_path / "inputs.pt")
    code = "\n".join([
        "import sys, torch",
        f"sys.path.insert(0, {ref!r})",
        "import spatial_vae.models as RM",
        f"p = torch.load({str(tmp_path / 'run_generator_epoch07.sav')!r}, weights_only=False)",
        f"q = torch.load({str(tmp_path / 'run_inference_epoch07.sav')!r}, weights_only=False)",
        "assert type(p) is RM.SpatialGenerator and type(q) is RM.InferenceNetwork and not p.training",
        f"d = torch.load({str(tmp_path / 'inputs.pt')!r})",
        "with torch.no_grad():",
        "    out = {'y_hat': p(d['x'], d['z']), 'zq': torch.cat(q(d['y']), 1)}",
        f"torch.save(out, {str(tmp_path / 'outputs.pt')!r})"])
    env = {k: v for k, v in os.environ.items() if k != "PYTHONPATH"}
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr[-2000:]
    out = torch.load(tmp_path / "outputs.pt")
    dec = O.decoder_params_from_state({k: v.detach() for k, v in p.state_dict().items()})
    enc = O.encoder_params_from_state({k: v.detach() for k, v in q.state_dict().items()})
    y_ref = O.decoder_forward(dec, x, z, "tanh", softplus=True)
    mu, ls = O.encoder_forward(enc, y, "tanh", True)
    assert torch.allclose(out["y_hat"], y_ref, rtol=1e-5, atol=1e-6)
    assert torch.allclose(out["zq"], torch.cat([mu, ls], 1), rtol=1e-5, atol=1e-6)


def test_driver_functions_keep_the_reference_signatures():
    """tests/golden/driver_signatures.json: inspect.signature of every function the reference's three driver scripts
    define (written from the reference in the build container).  The mirrors define the same names; parameter lists
    are identical or extend the reference's with trailing optional arguments (eps=..., argv=...)."""
    import importlib.util
    import inspect
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref = json.load(open(os.path.join(root, "tests", "golden", "driver_signatures.json")))
    mods = {}
    for key, sig in sorted(ref.items()):
        script, fn = key.split(".")
        if script not in mods:
            spec = importlib.util.spec_from_file_location("cli_sig_" + script,
                                                          os.path.join(root, "spatial-vae_b200", script + ".py"))
            mods[script] = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mods[script])
        obj = getattr(mods[script], fn, None)
        assert obj is not None, key
        mine = str(inspect.signature(obj))
        assert mine == sig or mine.startswith(sig[:-1]), (key, sig, mine)


def test_ingest_helpers_match_the_reference_fixture():
    """tests/golden/ingest.npz (written by the reference in the build container): Fourier downsampling, crop,
    background normalisation, the astigmatic 2-D CTF, and MRC files byte for byte (default header; make_header with an
    extended header and int16 voxels; a single section with cell dimensions), then parsed back."""
    import io
    import spatial_vae.ctf as C
    import spatial_vae.image as I
    import spatial_vae.mrc as M
    d = load_case("ingest")
    stack = d["stack"]
    for got, key in ((I.downsample(stack, factor=2), "down_factor2"), (I.downsample(stack, shape=(5, 7)), "down_shape_5x7"),
                     (I.downsample(stack[0], factor=1.5), "down_2d"), (I.normalize(stack), "norm_default"),
                     (I.normalize(stack, radius=3), "norm_r3")):
        assert got.dtype == d[key].dtype and got.shape == d[key].shape, key
        np.testing.assert_allclose(got, d[key], rtol=1e-5, atol=1e-6, err_msg=key)
    assert np.array_equal(I.crop(stack, 8), d["crop8"])
    np.testing.assert_allclose(C.compute_2d_ctf(d["freqs"], 21000.0, 18000.0, 0.6, 300.0, 2.7, 0.1, 120.0),
                               d["ctf_astig"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(C.compute_2d_ctf(d["freqs"], 15000.0, 15000.0, 0.0, 200.0, 2.0, 0.07), d["ctf_no_b"],
                               rtol=1e-10, atol=1e-12)

    buf = io.BytesIO()
    M.write(buf, stack)
    assert buf.getvalue() == d["mrc_default"].tobytes()
    vol = d["vol_int16"]
    hdr = M.make_header(vol.shape, (6.0, 4.0, 2.0), (90.0, 90.0, 90.0), mz=2, dtype=np.int16, exthd_size=16)
    buf = io.BytesIO()
    M.write(buf, vol, header=hdr, extended_header=bytes(range(16)))
    assert buf.getvalue() == d["mrc_int16_ext"].tobytes()
    buf = io.BytesIO()
    M.write(buf, stack[:1], ax=25.0, ay=30.0, az=1.0, alpha=90.0, beta=90.0, gamma=90.0)
    assert buf.getvalue() == d["mrc_single_cell"].tobytes()

    arr, h, ext = M.parse(d["mrc_int16_ext"].tobytes())
    assert np.array_equal(arr, vol) and arr.dtype == np.int16 and ext == bytes(range(16))
    assert (h.nx, h.ny, h.nz, h.mode, h.mz, h.next, h.mapc, h.mapr, h.maps) == (6, 4, 2, 1, 2, 16, 1, 2, 3)
    assert (h.xlen, h.ylen, h.zlen, h.amin, h.amax, h.amean, h.rms) == (6.0, 4.0, 2.0, 0.0, -1.0, -2.0, -1.0)
    arr, h, _ = M.parse(d["mrc_single_cell"].tobytes())
    assert arr.shape == stack.shape[1:] and np.array_equal(arr, stack[0]) and (h.xlen, h.ylen) == (25.0, 30.0)
    assert len(M.MRCHeader._fields) == 49 and M.header_struct.size == 1024
    assert [M.get_mode(np.dtype(t)) for t in ("int8", "int16", "float32", "complex64", "uint16")] == d["modes"].tolist()
    assert M.get_mode(np.dtype("2h")) == 3 and M.get_mode(np.dtype("3B")) == 16
