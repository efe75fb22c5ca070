"""CPU check of the first-layer arithmetic the CUDA kernels compile (spatial-vae_b200/csrc/first_layer.cuh).

The header is built with g++ (tests/csrc/first_layer_host.cpp wraps it in plain loops) and its forward value,
feature moments and closed-form (d theta, d t, d x) are compared with the oracle's decoder_forward + autograd for
every combination of --expand-coords / --bilinear (reference models.py:99-121).  The GEMM-shaped pieces around it
(W_eff, dWc, dWb, dz) are restated in numpy exactly as api.cu schedules them."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import svae_oracle as O

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def fl(tmp_path_factory):
    out = tmp_path_factory.mktemp("fl") / "first_layer_host.so"
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-o", str(out), str(ROOT / "tests/csrc/first_layer_host.cpp")],
                   check=True)
    lib = C.CDLL(str(out))
    for name in ("fl_forward", "fl_moments", "fl_latent", "fl_row_grad"):
        getattr(lib, name).restype = None
    return lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("expand,bilinear", [(False, False), (True, False), (False, True), (True, True)])
def test_first_layer_forward_and_backward(fl, expand, bilinear):
    torch.manual_seed(3)
    B, n, H, Z = 3, 6, 16, 4
    P, F = n * n, 5 if expand else 2
    grid = O.make_grid(n, n)
    theta = (torch.randn(B) * 0.7).requires_grad_()
    dx = (torch.randn(B, 2) * 0.1).requires_grad_()
    z = torch.randn(B, Z).requires_grad_()
    dec = {"coord_w": (torch.randn(H, F) * 0.5).requires_grad_(), "coord_b": (torch.randn(H) * 0.1).requires_grad_(),
           "latent_w": (torch.randn(H, Z) * 0.3).requires_grad_(), "hidden": [],
           "out_w": torch.randn(1, H) * 0.3, "out_b": torch.zeros(1)}
    if bilinear:
        dec["bilinear_w"] = (torch.randn(H, F, Z) * 0.2).requires_grad_()
    x = O.transform_coords(grid, theta, dx)
    x.retain_grad()
    y_hat = O.decoder_forward(dec, x, z, "tanh")
    wgt = torch.randn_like(y_hat)
    loss = (y_hat * wgt).sum()
    leaves = [theta, dx, z, dec["coord_w"], dec["coord_b"], dec["latent_w"]] + ([dec["bilinear_w"]] if bilinear else [])
    ref = dict(zip(["theta", "dx", "z", "coord_w", "coord_b", "latent_w", "bilinear_w"],
                   torch.autograd.grad(loss, leaves + [x], retain_graph=True)[:len(leaves)]))
    g_x_ref = torch.autograd.grad(loss, x, retain_graph=True)[0]

    # ---- what the library does -----------------------------------------------------------------------
    with torch.no_grad():
        img = torch.stack([torch.cos(theta), torch.sin(theta), dx[:, 0], dx[:, 1]], 1).contiguous().numpy()
        hz = (z @ dec["latent_w"].t() + dec["coord_b"]).contiguous().numpy()          # latent_projection
        if bilinear:                                                                     # W_eff (B, H*F) = z Wb^T + Wc
            w = (z @ dec["bilinear_w"].reshape(H * F, Z).t() + dec["coord_w"].reshape(-1)).contiguous().numpy()
            sb = H * F
        else:
            w, sb = dec["coord_w"].detach().contiguous().numpy(), 0
    g = grid.contiguous().numpy()
    pre = np.empty((B, P, H), np.float32)
    fl.fl_forward(F, B, P, H, _p(g), _p(img), _p(w), C.c_long(sb), C.c_long(F), C.c_long(1), _p(hz), _p(pre))

    # forward: tanh(pre) through the rest of the oracle decoder must reproduce y_hat
    h0 = torch.tanh(torch.from_numpy(pre)).requires_grad_()
    y2 = torch.sigmoid(h0 @ dec["out_w"].t() + dec["out_b"])
    assert torch.allclose(y2, y_hat.detach(), rtol=1e-5, atol=1e-6)
    d0 = torch.autograd.grad((y2 * wgt).sum(), h0)[0] * (1 - h0.detach() ** 2)           # delta0 = dL/d pre
    d0 = d0.contiguous().numpy()

    T = np.empty((B, F + 1, H), np.float32)
    fl.fl_moments(F, B, P, H, _p(g), _p(img), _p(d0), _p(T))
    lat = np.empty((B, 3), np.float32)
    fl.fl_latent(F, B, H, _p(w), C.c_long(sb), C.c_long(F), C.c_long(1), _p(T), _p(img), _p(lat))

    def close(a, b, what):
        a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
        assert np.allclose(a, b, rtol=2e-4, atol=2e-5 * max(1.0, np.abs(b).max())), (what, np.abs(a - b).max())

    close(lat[:, 0], ref["theta"], "d theta")
    close(lat[:, 1:], ref["dx"], "d dx")
    close(T[:, 0].sum(0), ref["coord_b"], "d coord_b")
    close(T[:, 1:].sum(0).T, ref["coord_w"], "d coord_w")                               # dWc[n,i] = sum_b T[b,1+i,n]
    close(T[:, 0].T @ z.detach().numpy(), ref["latent_w"], "d latent_w")
    dz = T[:, 0] @ dec["latent_w"].detach().numpy()
    if bilinear:
        wb = dec["bilinear_w"].detach().numpy()
        close(np.einsum("bin,bj->nij", T[:, 1:], z.detach().numpy()), ref["bilinear_w"], "d bilinear_w")
        dz = dz + np.einsum("bin,nij->bj", T[:, 1:], wb)
    close(dz, ref["z"], "d z")

    # module path: gradient w.r.t. explicit coordinates
    xe = x.detach().contiguous().numpy()
    gx = np.empty((B, P, 2), np.float32)
    fl.fl_row_grad(F, B, P, H, _p(xe), _p(w), C.c_long(sb), C.c_long(F), C.c_long(1), _p(d0), _p(gx))
    close(gx, g_x_ref, "d x")
