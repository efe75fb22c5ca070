"""Image-stack helpers of the particle pipeline (host preprocessing; same functions as reference
spatial_vae/image.py: downsample, crop, normalize)."""
import numpy as np


def downsample(x, factor=1, shape=None):
    """Fourier-domain downsampling of the last two axes: keep the lowest frequencies that fit the new shape (the
    half-spectrum of a real FFT: first m//2 and last m//2 rows, n//2+1 columns) and rescale by the pixel-count ratio
    (reference image.py:6-29)."""
    rows, cols = x.shape[-2:]
    if shape is None:
        shape = (int(rows / factor), int(cols / factor))
    m, n = shape
    spectrum = np.fft.rfft2(x)
    kept = np.concatenate([spectrum[..., :m // 2, :n // 2 + 1], spectrum[..., -m // 2:, :n // 2 + 1]], axis=-2)
    kept = kept * ((m * n) / (rows * cols))
    return np.fft.irfft2(kept, s=shape).astype(x.dtype)


def crop(stack, size):
    """Centre crop of the last two axes to size x size (reference image.py:32-44)."""
    n, m = stack.shape[-2:]
    top, left = (n - size) // 2, (m - size) // 2
    return stack[..., top:top + size, left:left + size]


def normalize(stack, radius=None):
    """Per image: subtract the mean and divide by the standard deviation of the BACKGROUND, the pixels at distance
    >= radius (default min(n, m) / 2) from (n/2, m/2) (reference image.py:47-63)."""
    n, m = stack.shape[-2:]
    if radius is None:
        radius = min(n, m) / 2
    yy, xx = np.ogrid[:n, :m]
    background = np.sqrt((n / 2 - yy) ** 2 + (m / 2 - xx) ** 2) >= radius
    pixels = stack[:, background]                               # (N, background pixels)
    mu = pixels.mean(axis=1).reshape(-1, 1, 1)
    sd = pixels.std(axis=1).reshape(-1, 1, 1)
    return ((stack - mu) / sd).astype(stack.dtype)
