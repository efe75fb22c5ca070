"""The three command lines on real input FILES (.mrcs stack + CTF table, .npy stacks) and with --save-prefix, run on
the CPU against tests/simt_emu: the ingest side of the path (mrc.parse, ctf_filter, normalise,
crop, mask) feeding the fused step, and the checkpoint side (whole-module .sav pickles the reference can read).
The GPU suite runs the same command lines on --synthetic data only."""
import contextlib
import glob
import importlib.util
import io
import math
import os

import numpy as np
import pytest
import torch

from tests import emu_backend

PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "spatial-vae_b200")


def _script(name):
    spec = importlib.util.spec_from_file_location("cli_files_" + name, os.path.join(PKG, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _run(mod, argv):
    with contextlib.redirect_stdout(io.StringIO()) as buf:
        mod.main(argv)
    lines = [l for l in buf.getvalue().splitlines() if "\t" in l]
    rows = [l.split("\t") for l in lines[1:]]
    return lines[0].split("\t"), rows


@pytest.fixture()
def emu(monkeypatch):
    emu_backend.install_all(monkeypatch)


def test_particles_cli_on_mrc_stack_with_ctf_table(emu, tmp_path):
    import spatial_vae.mrc as M
    rng = np.random.default_rng(0)
    n_train, n_test, n = 24, 8, 14
    for name, count in (("train", n_train), ("test", n_test)):
        with open(tmp_path / f"{name}.mrcs", "wb") as f:
            M.write(f, (rng.standard_normal((count, n, n)) * 2 + 5).astype(np.float32))
        # defocus cs voltage apix bfactor ampcont dfdiff dfang (reference spatial_vae/ctf.py:27-30)
        table = np.stack([rng.uniform(1, 3, count), np.full(count, 2.7), np.full(count, 300.0), np.full(count, 2.5),
                          np.full(count, 100.0), np.full(count, 10.0), np.zeros(count), rng.uniform(0, 180, count)], 1)
        np.savetxt(tmp_path / f"{name}_ctf.txt", table, fmt="%.6f")
    mod = _script("train_particles")
    header, rows = _run(mod, [str(tmp_path / "train.mrcs"), str(tmp_path / "test.mrcs"),
                              "--ctf-train", str(tmp_path / "train_ctf.txt"), "--ctf-test", str(tmp_path / "test_ctf.txt"),
                              "--normalize", "--crop", "12", "--mask", "--num-epochs", "2", "--minibatch-size", "10",
                              "--p-hidden-dim", "32", "--q-hidden-dim", "32", "--seed", "0", "--precision", "parity"])
    assert header[0] == "Epoch" and len(header) == 5 and len(rows) == 4
    assert all(math.isfinite(float(r[-3])) for r in rows)


def test_mnist_cli_on_npy_files_writes_checkpoints(emu, tmp_path, monkeypatch):
    rng = np.random.default_rng(1)
    os.makedirs(tmp_path / "data" / "mnist_rotated")
    for name, count in (("train", 60), ("test", 20)):
        img = ((rng.random((count, 10, 10)) > 0.8) * rng.random((count, 10, 10)) * 255).astype(np.uint8)
        np.save(tmp_path / "data" / "mnist_rotated" / f"images_{name}.npy", img)
    monkeypatch.chdir(tmp_path)
    mod = _script("train_mnist")
    prefix = "run"        # a label, not a path: outputs_<prefix>/trained/<prefix>_*_epochN.sav (misc_tools.py:57-99)
    header, rows = _run(mod, ["--dataset", "mnist-rotated", "-z", "3", "--num_epochs", "2", "--minibatch_size", "25",
                              "--p_hidden_dim", "32", "--q_hidden_dim", "32", "--save_prefix", prefix,
                              "--save_interval", "1", "--yes", "--seed", "0", "--precision", "parity"])
    assert header[0] == "Epoch" and len(rows) == 4
    vals = [float(r[-3]) for r in rows]
    assert all(math.isfinite(v) for v in vals)
    out = tmp_path / "outputs_run"
    savs = sorted(glob.glob(str(out / "trained" / "run_*_epoch*.sav")))
    assert [os.path.basename(f) for f in savs] == ["run_generator_epoch2.sav", "run_inference_epoch2.sav"]
    # (the mnist driver saves once, after the last epoch: reference train_mnist.py:448-451)
    assert (out / "command.txt").exists() and (out / "train.txt").exists() and (out / "val.txt").exists()
    gen = [s for s in savs if "generator" in s][-1]
    p = torch.load(gen, weights_only=False)
    assert type(p).__name__ == "SpatialGenerator" and type(p).__module__ == "spatial_vae.models"
    assert p.coord_linear.weight.shape == (32, 2) and p.latent_linear.weight.shape == (32, 3)


def test_galaxy_cli_on_npy_files(emu, tmp_path):
    rng = np.random.default_rng(2)
    np.save(tmp_path / "train.npy", rng.random((20, 8, 8, 3)).astype(np.float32))
    np.save(tmp_path / "test.npy", rng.random((8, 8, 8, 3)).astype(np.float32))
    mod = _script("train_galaxy")
    header, rows = _run(mod, [str(tmp_path / "train.npy"), str(tmp_path / "test.npy"), "-z", "4", "--num_epochs", "2",
                              "--minibatch_size", "10", "--p_hidden_dim", "32", "--p_num_layers", "3",
                              "--q_hidden_dim", "48", "--augment_rotation", "--seed", "0", "--precision", "parity"])
    assert header[0] == "Epoch" and len(rows) == 4
    assert all(math.isfinite(float(r[-3])) for r in rows)
