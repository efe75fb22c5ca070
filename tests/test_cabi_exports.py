"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/svae_b200.h declares (no compute calls without a GPU), and the host mirror of the
reference module API keeps names, keys and error behaviour."""
import contextlib
import io
import os
import re

import pytest
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as G
    G.build()
    import spatial_vae._lib as L
    return L


def test_library_exports_every_declared_symbol(built):
    hdr = open(os.path.join(ROOT, "include", "svae_b200.h")).read()
    declared = set(re.findall(r"^\s*(?:int|unsigned long long)\s+(svae_\w+)\s*\(", hdr, flags=re.M))
    assert declared, "no declarations parsed"
    assert declared == set(built.EXPORTS)
    for name in declared:
        assert hasattr(built.lib, name), name
    assert built.lib.svae_version() >= 100


def test_struct_layouts_match_header(built, tmp_path):
    """sizeof / offsetof of every struct as gcc lays out include/svae_b200.h == the ctypes mirror in _lib.py."""
    import ctypes as C
    import subprocess
    structs = {"SvaeShape": built.SvaeShape, "SvaeConfig": built.SvaeConfig,
               "SvaeDecoderParams": built.SvaeDecoderParams, "SvaeEncoderParams": built.SvaeEncoderParams,
               "SvaeStepInputs": built.SvaeStepInputs, "SvaeStepOutputs": built.SvaeStepOutputs}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "svae_b200.h"', 'int main(void) {']
    for name, cls in structs.items():
        lines.append(f'printf("{name} %zu\\n", sizeof({name}));')
        for field, _ in cls._fields_:
            lines.append(f'printf("{name}.{field} %zu\\n", offsetof({name}, {field}));')
    lines += ['return 0; }']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for name, cls in structs.items():
        assert int(got[name]) == C.sizeof(cls), name
        for field, _ in cls._fields_:
            assert int(got[f"{name}.{field}"]) == getattr(cls, field).offset, (name, field)
    # the header has no field the mirror lacks: count the members gcc sees through the struct sizes above and
    # through the declarations themselves
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "svae_b200.h")).read(), flags=re.S)
    for name, cls in structs.items():
        body = re.search(r"typedef struct \{([^}]*)\} %s;" % name, hdr).group(1)
        members = [m for m in body.split(";") if m.strip()]
        assert len(members) == len(cls._fields_), (name, members)


def test_invalid_arguments_return_error_codes_without_a_gpu(built):
    import ctypes as C
    s, c = built.SvaeShape(), built.SvaeConfig()
    s.B, s.P, s.H, s.L, s.C, s.Z, s.I = 4, 16, 32, 2, 5, 2, 5
    n = C.c_size_t(0)
    rc = built.lib.svae_workspace_bytes(C.byref(s), C.byref(c), C.byref(n))
    assert rc == -1 and "n_out" in built.last_error()
    s.C = 2
    c.rotate, c.translate, c.likelihood = 1, 1, built.LIK_GAUSS_FITNOISE
    s.k_ctf = 5
    rc = built.lib.svae_workspace_bytes(C.byref(s), C.byref(c), C.byref(n))
    assert rc == -1 and "fit-noise" in built.last_error()      # reference crashes on this combination
    s.k_ctf = 0
    assert built.lib.svae_workspace_bytes(C.byref(s), C.byref(c), C.byref(n)) == 0 and n.value > 0


def test_module_api_mirrors_reference(built):
    import spatial_vae.models as M
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        p = M.SpatialGenerator(3, 16, n_out=2, num_layers=3, activation=nn.Tanh)
        q = M.InferenceNetwork(30, 6, 12, num_layers=2, activation=nn.LeakyReLU)
    assert "SpatialGenerator" in buf.getvalue() and "InferenceNetwork" in buf.getvalue()   # ctor prints itself
    assert list(p.state_dict()) == ["coord_linear.weight", "coord_linear.bias", "latent_linear.weight",
                                    "layers.1.weight", "layers.1.bias", "layers.3.weight", "layers.3.bias",
                                    "layers.5.weight", "layers.5.bias"]
    assert list(q.state_dict()) == ["layers.0.weight", "layers.0.bias", "layers.2.weight", "layers.2.bias",
                                    "layers.4.weight", "layers.4.bias"]
    assert isinstance(p.layers[-1], nn.Sigmoid) and p.layers[5].out_features == 2
    # no CPU fallback: CPU tensors fail loudly
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        p(torch.zeros(2, 5, 2), torch.zeros(2, 3))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        q(torch.zeros(2, 30))


def test_same_seed_gives_reference_parameter_init(built):
    """Parameters are created in the reference's order (coord, latent, hidden..., out), so a
    golden fixture's seed reproduces its initial state_dict."""
    import numpy as np
    import spatial_vae.models as M
    from tests.helpers import load_case
    d = load_case("mnist_rt")
    torch.manual_seed(1)
    with contextlib.redirect_stdout(io.StringIO()):
        p = M.SpatialGenerator(3, 16, n_out=1, num_layers=2, activation=nn.Tanh)
        q = M.InferenceNetwork(30, 6, 12, num_layers=2, activation=nn.Tanh)
    for k, v in p.state_dict().items():
        np.testing.assert_array_equal(v.numpy(), d["p." + k])
    for k, v in q.state_dict().items():
        np.testing.assert_array_equal(v.numpy(), d["q." + k])


def test_option_networks_have_no_cpu_path_either(built):
    """resid / expand_coords / bilinear run through option_kernels.cu by default (their GPU parity tests are green:
    tests/test_gpu_zz_options.py); like everything else they refuse CPU tensors instead of falling back, and a
    standalone ResidLinear goes through the library as well."""
    import spatial_vae.models as M
    with contextlib.redirect_stdout(io.StringIO()):
        nets = [M.SpatialGenerator(3, 16, resid=True, num_layers=2), M.SpatialGenerator(3, 16, expand_coords=True),
                M.SpatialGenerator(3, 16, bilinear=True)]
        q = M.InferenceNetwork(8, 3, 16, num_layers=2, resid=True)
    assert [k for k, _ in nets[2].named_parameters()][:4] == ["coord_linear.weight", "coord_linear.bias",
                                                              "latent_linear.weight", "bilinear.weight"]
    assert "layers.1.linear.weight" in dict(nets[0].named_parameters())
    for p in nets:
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            p(torch.zeros(1, 4, 2), torch.zeros(1, 3))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        q(torch.zeros(2, 8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        M.ResidLinear(8, 8)(torch.zeros(2, 8))


def test_decoder_tensor_layout_follows_parameter_order(built):
    import spatial_vae.models as M
    import spatial_vae.functional as SF
    with contextlib.redirect_stdout(io.StringIO()):
        p = M.SpatialGenerator(3, 16, n_out=2, num_layers=3, resid=True, expand_coords=True, bilinear=True)
    dec = SF.decoder_tensors_of(p)
    assert [t.data_ptr() for t in dec.flat()] == [t.data_ptr() for t in p.parameters()]
    assert dec.layout() == (True, 2, True) and tuple(dec.bilinear_w.shape) == (16, 5, 3)
    again = SF.DecoderTensors.from_flat(dec.flat(), *dec.layout())
    assert [t.data_ptr() for t in again.flat()] == [t.data_ptr() for t in dec.flat()]
    with contextlib.redirect_stdout(io.StringIO()):
        q = M.InferenceNetwork(8, 3, 16, num_layers=3, resid=True)
    assert [t.data_ptr() for pr in SF.encoder_pairs_of(q) for t in pr] == [t.data_ptr() for t in q.parameters()]


def test_rotation_matrices_match_pillow_recipe(built):
    """svae_rotation_matrices is host code: compare with the PIL/Image.py recipe restated in the oracle."""
    import ctypes as C
    import numpy as np
    from oracle.pillow_rotate import rotate_matrix
    rng = np.random.default_rng(0)
    ang = np.concatenate([rng.uniform(-720, 720, 200), [0.0, 90.0, 180.0, 270.0, 360.0, -90.0]])
    B = len(ang)
    mats = np.zeros((B, 6))
    modes = np.zeros(B, dtype=np.int32)
    rc = built.lib.svae_rotation_matrices(ang.ctypes.data, B, 40, 40, mats.ctypes.data, modes.ctypes.data)
    assert rc == 0
    for b in range(B):
        a = ang[b] % 360.0
        expect_mode = {0.0: 1, 180.0: 2, 90.0: 3, 270.0: 4}.get(a, 0)
        assert modes[b] == expect_mode
        assert mats[b].tolist() == rotate_matrix(ang[b], 40, 40), (ang[b], mats[b], rotate_matrix(ang[b], 40, 40))


def test_product_sources_keep_to_the_rules():
    """Static guards: the product never imports the oracle or the test emulator, never names the batched-memcpy runtime
    entry points this environment forbids, and the library sources contain no CPU fallback switch."""
    import glob
    banned_api = ["cudaMemcpy" + "BatchAsync", "cudaMemcpy3D" + "BatchAsync", "cuMemcpy" + "BatchAsync",
                  "cuMemcpy3D" + "BatchAsync"]
    product = glob.glob(os.path.join(ROOT, "spatial-vae_b200", "**", "*.py"), recursive=True) + \
        glob.glob(os.path.join(ROOT, "spatial-vae_b200", "csrc", "*.cu*")) + [os.path.join(ROOT, "include", "svae_b200.h")]
    assert len(product) > 15
    for path in product:
        text = open(path).read()
        for name in banned_api:
            assert name not in text, (path, name)
        if path.endswith(".py"):
            assert "import oracle" not in text and "from oracle" not in text, path
            assert "simt_emu" not in text and "emu_backend" not in text, path
    bench = open(os.path.join(ROOT, "bench.py")).read()
    main_src = bench[bench.index("def main():"):]
    assert "oracle" not in main_src.split("cpu_reference_steps(c, sb")[0].replace("oracle port", ""), \
        "the measured arm of bench.py must not touch oracle/"
