"""CTF parameter tables -> per-particle real-space kernels (host precompute, numpy).

Replaces reference spatial_vae/ctf.py:7-56 for the drop-in CLI.  This is one-off data preparation,
not the hot path: the hot path consumes the (N, k, k) fp32 kernels this module produces.  Quirks kept:
`defocus` is used for both defocus axes (reference ctf.py:45-46, so `dfdiff`/`dfang` have no effect)
and the kernel is -fftshift(ifft2(CTF)).real (ctf.py:54).
"""
import numpy as np

CTF_COLUMNS = ['defocus', 'cs', 'voltage', 'apix', 'bfactor', 'ampcont', 'dfdiff', 'dfang']


def parse_ctf(f):
    """Whitespace separated table (path or file object), one row per particle, 8 columns (reference ctf.py:27-30)."""
    import pandas as pd
    table = pd.read_csv(f, sep=r'\s+', header=None)
    table.columns = CTF_COLUMNS
    return table


def electron_wavelength(kilovolts):
    v = np.asarray(kilovolts, dtype=np.float64) * 1e3
    return 12.2639 / np.sqrt(v + 0.97845e-6 * v * v)


def compute_2d_ctf(freqs, dfu, dfv, dfang, volt, cs, w, bfactor=None):
    """CTF at the spatial frequencies freqs (K, 2) in 1/Angstrom: defocus dfu / dfv (Angstrom) along / across the
    astigmatism axis dfang (radians), voltage in kV, spherical aberration cs in mm, amplitude contrast w, optional
    B-factor envelope (reference ctf.py:7-24)."""
    freqs = np.asarray(freqs)
    fx, fy = freqs[:, 0], freqs[:, 1]
    s2 = fx * fx + fy * fy
    lam = electron_wavelength(volt)
    defocus = 0.5 * (dfu + dfv + (dfu - dfv) * np.cos(2 * (np.arctan2(fy, fx) - dfang)))
    gamma = 2 * np.pi * (-0.5 * defocus * lam * s2 + 0.25 * (cs * 1e7) * lam ** 3 * s2 ** 2)
    ctf = np.sqrt(1 - w ** 2) * np.sin(gamma) - w * np.cos(gamma)
    if bfactor is not None:
        ctf = ctf * np.exp(-bfactor / 4 * s2)
    return ctf.astype(freqs.dtype)


def ctf_filter(ctf_params, n, m, scale=1):
    """(len(table), n, m) float32 real-space kernels, all particles at once."""
    col = lambda name: np.asarray(ctf_params[name], dtype=np.float64)
    fy, fx = np.meshgrid(np.fft.fftfreq(n), np.fft.fftfreq(m), indexing='ij')
    apix = col('apix')[:, None, None] * scale
    s2 = (fy[None] / apix) ** 2 + (fx[None] / apix) ** 2
    lam = electron_wavelength(col('voltage'))[:, None, None]
    cs = col('cs')[:, None, None] * 1e7
    df = col('defocus')[:, None, None] * 1e4
    w = col('ampcont')[:, None, None] / 100.0
    gamma = 2 * np.pi * (-0.5 * df * lam * s2 + 0.25 * cs * lam ** 3 * s2 ** 2)
    ctf = np.sqrt(1 - w ** 2) * np.sin(gamma) - w * np.cos(gamma)
    ctf = ctf * np.exp(-col('bfactor')[:, None, None] / 4 * s2)
    kernels = -np.fft.fftshift(np.fft.ifft2(ctf, axes=(-2, -1)), axes=(-2, -1)).real
    return kernels.astype(np.float32)
