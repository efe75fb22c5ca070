"""CPU restatement of Pillow's Image.rotate(angle, resample=BICUBIC) as the reference drivers use it for
--augment-rotation (reference train_particles.py:39-43 on float32 "F" images; train_galaxy.py:47-54 on uint8 RGB).

TEST INFRASTRUCTURE ONLY (see oracle/svae_oracle.py).  The arithmetic lives in a third-party dependency of
the reference, Pillow (pinned Pillow~=8.2.0 in requirements.txt:5; 12.2.0 in this image): PIL/Image.py builds the
inverse affine matrix, libImaging/Geometry.c (ImagingGenericTransform + bicubic_filter*) resamples in double
precision with the a = -1 cubic, edge replication inside the 4x4 window and zero fill outside the image.
Pinned by tests/test_oracle_golden.py against the installed Pillow on random images and angles (bit exact).
"""
import math

import numpy as np


def rotate_matrix(angle_deg, w, h):
    """Inverse (destination -> source) affine of Image.rotate about the image centre (PIL/Image.py)."""
    a = -math.radians(angle_deg % 360.0)
    m = [round(math.cos(a), 15), round(math.sin(a), 15), 0.0, round(-math.sin(a), 15), round(math.cos(a), 15), 0.0]
    cx, cy = w / 2, h / 2
    m[2] = m[0] * -cx + m[1] * -cy + m[2] + cx
    m[5] = m[3] * -cx + m[4] * -cy + m[5] + cy
    return m


def _cubic(v1, v2, v3, v4, d):
    """Geometry.c BICUBIC macro (a = -1 cubic).  The coefficient sums are evaluated in the type of the samples:
    float32 for the horizontal pass over an "F" image (C float arithmetic), double for the vertical pass and for
    integer samples; the Horner evaluation is always in double."""
    p1 = v2
    p2 = -v1 + v3
    p3 = 2 * (v1 - v2) + v3 - v4
    p4 = -v1 + v2 - v3 + v4
    d = np.asarray(d, dtype=np.float64)
    return p1.astype(np.float64) + d * (p2.astype(np.float64) + d * (p3.astype(np.float64) + d * p4.astype(np.float64)))


def rotate_bicubic(img, angle_deg):
    """img (h, w) float32 or (h, w, C) uint8 -> rotated image of the same shape/dtype."""
    angle = angle_deg % 360.0
    if angle == 0:
        return img.copy()
    h, w = img.shape[:2]
    if angle == 180:
        return img[::-1, ::-1].copy()
    if angle in (90, 270) and w == h:
        return np.rot90(img, 1 if angle == 90 else 3).copy()
    m = rotate_matrix(angle_deg, w, h)
    # horizontal pass arithmetic happens in the sample type (float32 for "F", exact integers for uint8)
    src = img.astype(np.float32) if img.dtype == np.float32 else img.astype(np.float64)
    ys, xs = np.meshgrid(np.arange(h) + 0.5, np.arange(w) + 0.5, indexing="ij")
    xin = m[0] * xs + m[1] * ys + m[2]
    yin = m[3] * xs + m[4] * ys + m[5]
    inside = (xin >= 0.0) & (xin < w) & (yin >= 0.0) & (yin < h)
    xin, yin = xin - 0.5, yin - 0.5
    x = np.floor(xin).astype(np.int64)
    y = np.floor(yin).astype(np.int64)
    dx, dy = xin - x, yin - y
    x, y = x - 1, y - 1
    cols = [np.clip(x + k, 0, w - 1) for k in range(4)]
    rows = []
    prev = None
    for k in range(4):
        yy = y + k
        if k == 0:
            r = np.clip(yy, 0, h - 1)
            val = _cubic(*(src[r, c] for c in cols), dx[..., None] if src.ndim == 3 else dx)
        else:
            ok = (yy >= 0) & (yy < h)
            r = np.clip(yy, 0, h - 1)
            val = _cubic(*(src[r, c] for c in cols), dx[..., None] if src.ndim == 3 else dx)
            val = np.where(ok[..., None] if src.ndim == 3 else ok, val, prev)
        rows.append(val)
        prev = val
    v = _cubic(rows[0], rows[1], rows[2], rows[3], dy[..., None] if src.ndim == 3 else dy)
    mask = inside[..., None] if src.ndim == 3 else inside
    if img.dtype == np.uint8:
        v = np.where(v <= 0.0, 0.0, np.where(v >= 255.0, 255.0, np.floor(v)))      # (UINT8) cast truncates
        return np.where(mask, v, 0.0).astype(np.uint8)
    return np.where(mask, v, 0.0).astype(np.float32)
